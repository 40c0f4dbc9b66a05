"""Mirror of the reference ``models/tokenizer.py`` (``SimpleTokenizer``, :63-151): the CLIP byte-level BPE tokenizer
that feeds ``CLIPEncoder.encode_prompt`` (models/vlm.py:67-69).  Same constructor argument (path of the gzipped merges
file), same ``encode`` / ``decode`` / ``__call__(texts, context_length=77)`` results: ``[n,77]`` int64 with
``<|startoftext|>`` ... ``<|endoftext|>`` and zero padding, truncated at the context length (:136-151); a single text
gives a 1-D tensor (:149-150).

Pure host string work (nothing here touches the GPU); written from the published algorithm, pinned against the
reference class on a committed synthetic merges file (tests/golden/tokenizer_*.json, oracle/make_golden_tokenizer.py)
because the real ``bpe_simple_vocab_16e6.txt.gz`` is not available offline.

Text cleaning: the reference runs ``ftfy.fix_text`` first (:52).  ``ftfy`` is used when importable; otherwise only
text on which ``fix_text`` is the identity is accepted (printable ASCII plus ordinary whitespace -- every Oryon
prompt) and anything else raises instead of silently tokenising differently.
"""
from __future__ import annotations

import gzip
import html
from typing import Dict, Iterable, List, Sequence, Tuple, Union

import regex
import torch

SOT, EOT = "<|startoftext|>", "<|endoftext|>"
END = "</w>"
_N_MERGES = 49152 - 256 - 2          # the released vocabulary: 48 894 merges (tokenizer.py:67)
_WORDS = regex.compile(
    r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)
_SPACES = regex.compile(r"\s+")


def byte_alphabet() -> List[str]:
    """One printable unicode character per byte value 0..255, indexed by the byte (tokenizer.py:17-37): bytes that are
    already printable latin-1 keep their code point, the other 68 are moved to U+0100 onwards in byte order."""
    keep = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    table, moved = [], 0
    for b in range(256):
        if b in keep:
            table.append(chr(b))
        else:
            table.append(chr(256 + moved))
            moved += 1
    return table


def vocabulary_order(alphabet: Sequence[str]) -> List[str]:
    """The order in which single characters enter the vocabulary (tokenizer.py:69): printable bytes first, then
    the moved ones -- i.e. the order of the reference's ``bytes_to_unicode().values()``."""
    printable = [c for c in alphabet if ord(c) < 256]
    return printable + [c for c in alphabet if ord(c) >= 256]


def _fix_text(text: str) -> str:
    try:
        import ftfy
    except ImportError:
        ftfy = None
    if ftfy is not None:
        return ftfy.fix_text(text)
    if all((0x20 <= ord(c) < 0x7F) or c in "\t\n" for c in text):
        return text
    raise RuntimeError("SimpleTokenizer: text with control or non-ASCII characters needs the `ftfy` package "
                       "(reference models/tokenizer.py:52); it is not installed")


def clean(text: str) -> str:
    """``whitespace_clean(basic_clean(text)).lower()`` (tokenizer.py:51-60, :124)."""
    text = html.unescape(html.unescape(_fix_text(text))).strip()
    return _SPACES.sub(" ", text).strip().lower()


class SimpleTokenizer:
    def __init__(self, bpe_path: str):
        with gzip.open(bpe_path) as fh:
            lines = fh.read().decode("utf-8").split("\n")
        merges: List[Tuple[str, ...]] = [tuple(line.split()) for line in lines[1:_N_MERGES + 1]]   # line 0 is the header
        self.alphabet = byte_alphabet()
        singles = vocabulary_order(self.alphabet)
        vocab = singles + [c + END for c in singles] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder: Dict[str, int] = {tok: i for i, tok in enumerate(vocab)}       # later duplicates win, like dict(zip(..))
        self.decoder: Dict[int, str] = {i: tok for tok, i in self.encoder.items()}
        self.bpe_ranks: Dict[Tuple[str, ...], int] = {m: r for r, m in enumerate(merges)}
        self.byte_of = {c: b for b, c in enumerate(self.alphabet)}
        self._memo: Dict[str, List[str]] = {SOT: [SOT], EOT: [EOT]}

    # ---- BPE ------------------------------------------------------------------------------------------
    def _split(self, token: str) -> List[str]:
        """Sub-word units of one pre-token: start from characters (the last one carries ``</w>``), repeatedly fuse
        every occurrence (left to right) of the adjacent pair with the lowest merge rank (tokenizer.py:82-121)."""
        hit = self._memo.get(token)
        if hit is not None:
            return hit
        parts = list(token[:-1]) + [token[-1] + END]
        ranks = self.bpe_ranks
        while len(parts) > 1:
            best, best_rank = None, None
            for pair in zip(parts, parts[1:]):
                r = ranks.get(pair)
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = pair, r
            if best is None:
                break
            fused, i, n = [], 0, len(parts)
            while i < n:
                if i + 1 < n and parts[i] == best[0] and parts[i + 1] == best[1]:
                    fused.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    fused.append(parts[i])
                    i += 1
            parts = fused
        self._memo[token] = parts
        return parts

    def bpe(self, token: str) -> str:
        return " ".join(self._split(token))

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for word in _WORDS.findall(clean(text)):
            mapped = "".join(self.alphabet[b] for b in word.encode("utf-8"))
            ids.extend(self.encoder[unit] for unit in self._split(mapped))
        return ids

    def decode(self, tokens: Iterable[int]) -> str:
        chars = "".join(self.decoder[int(t)] for t in tokens)
        return bytearray(self.byte_of[c] for c in chars).decode("utf-8", errors="replace").replace(END, " ")

    def __call__(self, texts: Union[str, Sequence[str]], context_length: int = 77) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        sot, eot = self.encoder[SOT], self.encoder[EOT]
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for row, text in zip(out, texts):
            ids = ([sot] + self.encode(text) + [eot])[:context_length]
            row[:len(ids)] = torch.tensor(ids, dtype=torch.long)
        return out[0] if len(texts) == 1 else out
