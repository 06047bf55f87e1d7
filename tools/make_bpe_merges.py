"""Derive the 48 894 BPE merge rules the CLIP tokenizer uses from the public OpenAI CLIP vocabulary file
(bpe_simple_vocab_16e6.txt.gz, MIT licence; 262 145 lines of which CLIP reads lines 1..48894) and write them as
clip_decontamination_b200/open_clip/clip_bpe_merges.txt.gz -- the data file open_clip/tokenizer.py ships with, so the
text cache (segmentor.py:157-174) can be built on a box without the reference checkout.

    python tools/make_bpe_merges.py [/path/to/bpe_simple_vocab_16e6.txt.gz]
"""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/open_clip/bpe_simple_vocab_16e6.txt.gz'
lines = gzip.open(src).read().decode('utf-8').split('\n')
merges = lines[1:49152 - 256 - 2 + 1]
assert len(merges) == 48894 and all(len(m.split()) == 2 for m in merges)
dst = os.path.join(ROOT, 'clip_decontamination_b200', 'open_clip', 'clip_bpe_merges.txt.gz')
with gzip.GzipFile(dst, 'wb', mtime=0) as f:
    f.write(('\n'.join(merges) + '\n').encode('utf-8'))
print(dst, os.path.getsize(dst), 'bytes,', len(merges), 'merges')
