"""CLIP byte-pair-encoding tokenizer (init-time text cache only, segmentor.py:157-174).

Own implementation of the published CLIP BPE scheme (lower-case, whitespace clean, GPT-2 style byte
-> unicode table, merges ranked by the vocabulary file, <start_of_text> / <end_of_text>, zero padding
to 77) -- behaviour contract: open_clip/tokenizer.py:83-85,250-257,270 of the reference.

The merge table is data: ``clip_bpe_merges.txt.gz`` next to this file holds the 48 894 merge rules CLIP reads from
the public OpenAI vocabulary (``bpe_simple_vocab_16e6.txt.gz`` lines 1..48894; derived by tools/make_bpe_merges.py).
``$CLIPSEG_BPE_VOCAB`` may point at either file instead.
"""
import gzip
import html
import os
from functools import lru_cache
from typing import List, Union

import regex as re
import torch

CONTEXT_LENGTH = 77
_VOCAB_NAME = 'bpe_simple_vocab_16e6.txt.gz'
_MERGES_NAME = 'clip_bpe_merges.txt.gz'
_N_MERGES = 49152 - 256 - 2


def find_bpe_vocab():
    here = os.path.dirname(os.path.abspath(__file__))
    for c in (os.environ.get('CLIPSEG_BPE_VOCAB', ''), os.path.join(here, _MERGES_NAME), os.path.join(here, _VOCAB_NAME)):
        if c and os.path.isfile(c):
            return c
    return None


@lru_cache()
def _byte_table():
    keep = list(range(ord('!'), ord('~') + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    chars, extra = keep[:], 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return {b: chr(c) for b, c in zip(keep, chars)}


class BPETokenizer:
    def __init__(self, vocab_path: str):
        self.byte_enc = _byte_table()
        lines = gzip.open(vocab_path).read().decode('utf-8').split('\n')
        if lines[0].startswith('"bpe_simple_vocab') or '#version' in lines[0]:      # the original OpenAI file: header line
            lines = lines[1:]
        merges = [tuple(m.split()) for m in lines[:_N_MERGES]]
        assert len(merges) == _N_MERGES and all(len(m) == 2 for m in merges), 'truncated BPE merge table'
        vocab = list(self.byte_enc.values())
        vocab = vocab + [v + '</w>' for v in vocab] + [''.join(m) for m in merges]
        vocab += ['<start_of_text>', '<end_of_text>']
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.sot, self.eot = self.encoder['<start_of_text>'], self.encoder['<end_of_text>']
        self.cache = {}
        self.pat = re.compile(r"<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+",
                              re.IGNORECASE)

    def _bpe(self, token: str) -> List[str]:
        if token in self.cache:
            return self.cache[token]
        word = list(token[:-1]) + [token[-1] + '</w>']
        while len(word) > 1:
            pairs = [(word[i], word[i + 1]) for i in range(len(word) - 1)]
            best = min(pairs, key=lambda p: self.rank.get(p, float('inf')))
            if best not in self.rank:
                break
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and (word[i], word[i + 1]) == best:
                    merged.append(word[i] + word[i + 1])
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        self.cache[token] = word
        return word

    def encode(self, text: str) -> List[int]:
        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r'\s+', ' ', text).strip().lower()
        ids = []
        for tok in re.findall(self.pat, text):
            tok = ''.join(self.byte_enc[b] for b in tok.encode('utf-8'))
            ids.extend(self.encoder[t] for t in self._bpe(tok))
        return ids

    def __call__(self, texts: Union[str, List[str]], context_length: int = CONTEXT_LENGTH) -> torch.LongTensor:
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), context_length, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.sot] + self.encode(t) + [self.eot]
            if len(ids) > context_length:
                ids = ids[:context_length]
                ids[-1] = self.eot
            out[i, :len(ids)] = torch.tensor(ids)
        return out


_tok = None


def tokenize(texts: Union[str, List[str]], context_length: int = CONTEXT_LENGTH) -> torch.LongTensor:
    """open_clip.tokenizer.tokenize (open_clip/tokenizer.py:270)."""
    global _tok
    if _tok is None:
        path = find_bpe_vocab()
        if path is None:
            raise RuntimeError(f'{_MERGES_NAME} not found next to {__file__} (set CLIPSEG_BPE_VOCAB)')
        _tok = BPETokenizer(path)
    return _tok(texts, context_length)
