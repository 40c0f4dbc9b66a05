"""Host-side mirrors of the reference's ``models/`` modules that sit on the inference path (tokenizer only; the
network itself lives in liboryon_b200.so, see oryon_b200/net.py)."""
