#!/bin/bash
set -u
bash tools/collect_round_profiles.sh r2_final 2>&1 | tail -45 | cut -c1-300
du -sh gpurun_out
