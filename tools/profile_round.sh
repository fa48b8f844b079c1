#!/bin/bash
# Round profile pass (run on the GPU box through gpurun):  bash tools/profile_round.sh <tag> [parts]
# parts (default: all) = any of: plain launches c2 c1 c4 c3
#   plain    plain bench line (never under a profiler)
#   launches ncu launch list of the same command
#   opt      `--set full` capture of the fused optimizer kernels (C5 tensors)
#   c2/c1/c4/c3  `--set full` captures of the hot kernels of C2 (hash / MLP64 / composite / march), C1 (tcgen05),
#                C4 (fused MLPs) and C3 (wide-input tcgen05 decoder)
# (no --import-source, and at most ~60 MB of reports per call: gpurun returns at most 64 MiB of gpurun_out/)
# Outputs land in gpurun_out/<tag>_*; profiles/summarize.py turns them into the committed summaries.
set -u
TAG=${1:-r1d}
PARTS=${2:-"plain launches c2 opt c1 c4 c3"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $PARTS " == *" $1 "* ]]; }
if has plain; then
  python bench.py > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || exit 1
fi
if has launches; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_ncu1.log 2>&1
fi
if has c2; then
  # 8 matching kernels per step (march mask / compact, hash fwd, decoder fwd, composite fwd / bwd, decoder bwd, hash bwd):
  # skip three steps, capture the fourth
  ncu --set full --clock-control none -k 'regex:k_hash|k_instant|k_composite|k_march' -s 24 -c 8 \
      -f -o $OUT/${TAG}_prof python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $OUT/${TAG}_ncu2.log 2>&1
fi
if has opt; then
  ncu --set full --clock-control none -k 'regex:k_opt_' -s 8 -c 2 \
      -f -o $OUT/${TAG}_prof_opt python bench.py --only c5_dualhash > $OUT/${TAG}_ncu6.log 2>&1
fi
if has c1; then
  ncu --set full --clock-control none -k 'regex:^k_mlp256$|k_wgrad256' -s 15 -c 3 \
      -f -o $OUT/${TAG}_prof_c1 python bench.py --only c1_vanilla > $OUT/${TAG}_ncu3.log 2>&1
fi
if has c4; then
  ncu --set full --clock-control none -k 'regex:k_fmlp|k_hash_bwd_input|k_wgrad256' -s 16 -c 4 \
      -f -o $OUT/${TAG}_prof_c4 python bench.py --only c4_instant_dnerf > $OUT/${TAG}_ncu4.log 2>&1
fi
if has c3; then
  ncu --set full --clock-control none -k 'regex:^k_mlp256$|k_nerf_dx' -s 9 -c 3 \
      -f -o $OUT/${TAG}_prof_c3 python bench.py --only c3_dnerf > $OUT/${TAG}_ncu5.log 2>&1
fi
ls -la $OUT; [ -f $OUT/${TAG}_plain.json ] && tail -c 300 $OUT/${TAG}_plain.json
