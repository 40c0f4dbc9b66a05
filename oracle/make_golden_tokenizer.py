"""Generate ``tests/golden/bpe_synth_vocab.txt.gz`` and ``tests/golden/tokenizer_synth.json`` by running the
UNMODIFIED reference ``models/tokenizer.py`` (loaded by file path from /root/reference).

TEST INFRASTRUCTURE.  Build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_tokenizer.py

The real CLIP merges file (``pretrained_models/bpe_simple_vocab_16e6.txt.gz``, models/vlm.py:23) is not available
offline, so the pin uses a synthetic merges file in the same wire format (header line, one ``left right`` merge per
line, gzip): a small byte-level BPE trained here on an Oryon-style prompt corpus.  ``ftfy`` is not installed; the
reference module is loaded with ``ftfy.fix_text`` stubbed to the identity, which is what ftfy does on the ASCII
prompts used below.
"""
import collections
import gzip
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REFERENCE_ROOT = os.environ.get("ORYON_REFERENCE_ROOT", "/root/reference")

TEMPLATES = ["a photo of a {}.", "a bad photo of the {}.", "a sculpture of a {}.", "a low resolution photo of the {}.",
             "a rendering of a {}.", "graffiti of a {}.", "a cropped photo of the {}.", "a bright photo of a {}.",
             "a close-up photo of a {}.", "a black and white photo of the {}.", "a 3d rendering of the {}.",
             "itap of my {}.", "a jpeg corrupted photo of a {}.", "the origami {}.", "a {} in a video game.",
             "a doodle of the {}.", "the plushie {}.", "art of the {}.", "a tattoo of the {}.", "a toy {}."]
OBJECTS = ["brown open laptop", "black camera", "white mug with a blue handle", "green plastic bottle", "red bowl",
           "small metal can", "yellow toy duck", "grey remote control", "wooden spoon", "children's book",
           "it's a 35mm lens", "object #7 (don't touch)", "orange &amp; white cat statue", "T-Shirt   XL",
           "mug", "x", "<|startoftext|> marker <|endoftext|>", "we'll see: they've got 2 cups, I'm sure; he'd say so"]
LONG = " ".join(["a very long description of a brown open laptop on a wooden table"] * 12)


def load_reference_tokenizer():
    sys.dont_write_bytecode = True
    if "ftfy" not in sys.modules:
        stub = types.ModuleType("ftfy")
        stub.fix_text = lambda text: text
        sys.modules["ftfy"] = stub
    spec = importlib.util.spec_from_file_location("ref_tokenizer", os.path.join(REFERENCE_ROOT, "models", "tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def train_merges(ref, corpus, n_merges):
    """Tiny byte-level BPE trainer (greedy most-frequent pair, ties by first occurrence) over the reference's own
    pre-tokenisation, producing merges in the released file's format."""
    enc = ref.bytes_to_unicode()
    import regex
    pat = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", regex.IGNORECASE)
    words = collections.Counter()
    for text in corpus:
        for tok in pat.findall(ref.whitespace_clean(ref.basic_clean(text)).lower()):
            sym = [enc[b] for b in tok.encode("utf-8")]
            sym[-1] += "</w>"
            words[tuple(sym)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), cnt = max(pairs.items(), key=lambda kv: kv[1])
        if cnt < 2:
            break
        merges.append((a, b))
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        words = new
    return merges


def main():
    ref = load_reference_tokenizer()
    corpus = [t.format(o) for t in TEMPLATES for o in OBJECTS[:12]]
    merges = train_merges(ref, corpus, 400)
    path = os.path.join(OUT, "bpe_synth_vocab.txt.gz")
    with open(path, "wb") as raw:   # mtime=0: reproducible bytes
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as fh:
            fh.write(('"bpe_synth_vocab: %d merges"\n' % len(merges) + "\n".join(" ".join(m) for m in merges) + "\n").encode("utf-8"))
    tok = ref.SimpleTokenizer(path)
    texts = [t.format(o) for o in OBJECTS for t in TEMPLATES[:6]] + OBJECTS + [LONG, "", "   ", "A  PHOTO\tOF\nA   MUG"]
    golden = {"n_merges": len(merges), "vocab_size": len(tok.encoder), "sot": tok.encoder["<|startoftext|>"],
              "eot": tok.encoder["<|endoftext|>"],
              "texts": texts, "encode": [tok.encode(t) for t in texts],
              "call77": tok(texts).tolist(), "single": tok(texts[0]).tolist(), "call16": tok(texts, context_length=16).tolist(),
              "decode": [tok.decode(tok.encode(t)) for t in texts],
              "bpe": {w: tok.bpe(w) for w in ["laptop", "photo", "a", "zzzzqq", "rendering", "'s"]}}
    with open(os.path.join(OUT, "tokenizer_synth.json"), "w") as fh:
        json.dump(golden, fh)
    print(f"{len(merges)} merges, vocab {len(tok.encoder)}, {len(texts)} texts -> {path}")


if __name__ == "__main__":
    main()
