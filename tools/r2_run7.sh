#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/collect_round_profiles.sh r2_final 2>&1 | tail -40
