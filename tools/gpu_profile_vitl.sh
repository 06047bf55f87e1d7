#!/bin/bash
# ncu evidence for BASELINE config 3 (ViT-L/14, LoveDA-shaped 1024^2, no upsampler): launch list of one step (2 tiles, 162 crops)
# and --set full rows of the attention / GEMM / LayerNorm kernels of a 3-block truncation.  Run through gpurun from the repo root.
set -u
OUT=gpurun_out
TAG=${1:-r02_vitl}
NCU="ncu --profile-from-start off --clock-control none"
W="--workload loveda1024_vitl --tiles 2"
timeout 120 python tools/profile_step.py $W > $OUT/prof_vitl_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/prof_vitl_plain.log; exit 1; }
timeout 200 $NCU --metrics gpu__time_duration.sum --csv --log-file $OUT/${TAG}_launches_step_t2.csv python tools/profile_step.py $W > $OUT/prof_vitl_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 120 python tools/profile_step.py $W --layers 3 > $OUT/prof_vitl_plainA.log 2>&1 && \
timeout 300 $NCU --set full -k 'regex:attention|gemm_bf16|layernorm' -f -o /tmp/${TAG}_full \
    python tools/profile_step.py $W --layers 3 > $OUT/prof_vitl_ncuA.log 2>&1; echo "full rc=$?"
ncu -i /tmp/${TAG}_full.ncu-rep --page raw --csv > /tmp/${TAG}_raw.csv 2>/dev/null && python tools/ncu_trim.py /tmp/${TAG}_raw.csv $OUT/${TAG}_ncu_full_3layers.csv
ls -la $OUT | tail -5
