#!/bin/bash
# one gpurun job: parity suite (all failures listed), then kernel A/B benches.  Usage: bash tools/gpu_job.sh <tag> [parts]
TAG=${1:-job}
PARTS=${2:-"tests hashv red"}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $PARTS " == *" $1 "* ]]; }
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
  cp $OUT/parity_maxima.json $OUT/${TAG}_parity_maxima.json 2>/dev/null
fi
if has smoke; then timeout 600 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; fi
if has hashv; then timeout 600 python tools/kbench.py hashv > $OUT/${TAG}_hashv.log 2>&1; fi
if has mlp64; then timeout 600 python tools/kbench.py mlp64 > $OUT/${TAG}_mlp64.log 2>&1; fi
if has itc; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "instant" > $OUT/${TAG}_itc.log 2>&1; fi
if has red; then timeout 300 python tools/kbench.py red > $OUT/${TAG}_red.log 2>&1; fi
if has bench; then timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; fi
if has benchc2; then timeout 900 python bench.py --no-extras > $OUT/${TAG}_benchc2.json 2> $OUT/${TAG}_benchc2.err; fi
ls -la $OUT | tail -5
tail -5 $OUT/${TAG}_tests.log 2>/dev/null
