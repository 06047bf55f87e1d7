"""Stage the UNMODIFIED reference for a GPU-box run: copies the handful of reference files the hot path imports from
/root/reference into the git-ignored ``baseline/_ref/``.  ``gpurun`` snapshots the working
tree, so the staged files travel to the B200 box, where ``oracle/ref_on_gpu.py`` (and ``bench.py --impl reference``
with CLIPSEG_REF_ARM=unmodified) import them through ``oracle/ref_harness.py`` exactly as the container does from
/root/reference.  Nothing staged here is tracked by git; nothing in the product path reads it.
TEST / MEASUREMENT INFRASTRUCTURE.   python -m oracle.stage_ref [--clean]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')
ITEMS = ['open_clip', 'simfeatup_dev/__init__.py', 'simfeatup_dev/upsamplers.py', 'prompts', 'configs',
         'segmentor.py', 'segearth_segmentor.py', 'outlier_suppression.py', 'similarity_enhancement.py', 'CTD.py',
         'self_attention_enhancement.py', 'cross_tile_fusion.py']


def main():
    if '--clean' in sys.argv:
        shutil.rmtree(DST, ignore_errors=True)
        print('removed', DST)
        return
    if not os.path.isdir(SRC):
        raise SystemExit(f'{SRC} not found: staging only works in the build container')
    os.makedirs(DST, exist_ok=True)
    n = 0
    for it in ITEMS:
        s, d = os.path.join(SRC, it), os.path.join(DST, it)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns('__pycache__', '*.pt', '*.ckpt'))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
        n += 1
    size = sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(DST) for f in fs)
    print(f'staged {n} items ({size / 1e6:.1f} MB) into {DST} (git-ignored)')


if __name__ == '__main__':
    main()
