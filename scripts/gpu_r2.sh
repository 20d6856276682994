#!/bin/bash
# Round-2 iteration gpurun: all GPU tests, the parity diagnostic (flip rates / distance errors per engine), one bench line.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2}; shift || true
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider ${PYTEST_ARGS:-} > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -15 $OUT/${TAG}_tests.log | cut -c1-300
if [ -z "${NO_PARITY:-}" ]; then echo "== parity"; timeout 900 python scripts/diag_parity.py --segments ${PARITY_SEGMENTS:-512} > $OUT/${TAG}_parity.json 2> $OUT/${TAG}_parity.err; echo "rc=$?"; tail -3 $OUT/${TAG}_parity.err; fi
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
