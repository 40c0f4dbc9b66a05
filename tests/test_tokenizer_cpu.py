"""a3 (host part): ``oryon_b200.models.tokenizer.SimpleTokenizer`` against outputs of the reference class
(models/tokenizer.py:63-151) recorded by oracle/make_golden_tokenizer.py on a synthetic merges file."""
import json
import os

import pytest
import torch

from oryon_b200.models.tokenizer import SimpleTokenizer, byte_alphabet, clean

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "tokenizer_synth.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def tok():
    return SimpleTokenizer(os.path.join(GOLDEN, "bpe_synth_vocab.txt.gz"))


def test_vocabulary_layout(tok, gold):
    assert len(tok.encoder) == gold["vocab_size"]
    assert tok.encoder["<|startoftext|>"] == gold["sot"] and tok.encoder["<|endoftext|>"] == gold["eot"]
    assert len(tok.bpe_ranks) == gold["n_merges"] + 1      # the trailing empty line of the file is a (never matching) merge
    alpha = byte_alphabet()
    assert len(set(alpha)) == 256 and alpha[ord("a")] == "a" and alpha[0] == chr(256) and alpha[0x20] == chr(256 + 32)


def test_encode_matches_reference(tok, gold):
    for text, ids in zip(gold["texts"], gold["encode"]):
        assert tok.encode(text) == ids, text


def test_bpe_units(tok, gold):
    for word, units in gold["bpe"].items():
        assert tok.bpe(word) == units


def test_call_shapes_padding_truncation(tok, gold):
    out = tok(gold["texts"])
    assert out.dtype == torch.long and out.shape == (len(gold["texts"]), 77)
    assert out.tolist() == gold["call77"]
    assert tok(gold["texts"], context_length=16).tolist() == gold["call16"]
    single = tok(gold["texts"][0])
    assert single.dim() == 1 and single.tolist() == gold["single"]
    # EOT is the largest id: vlm.py:81 locates it with argmax
    row = out[0]
    assert int(row.argmax()) == row.tolist().index(gold["eot"])


def test_decode_round_trip(tok, gold):
    for text, dec in zip(gold["texts"], gold["decode"]):
        assert tok.decode(tok.encode(text)) == dec


def test_clean_and_loud_failure():
    assert clean("  A &amp;amp; B\t\n c ") == "a & b c"
    try:
        import ftfy  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError):
            clean("café")
