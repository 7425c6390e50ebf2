#!/usr/bin/env bash
# First GPU call of the next round (run from the repo root through gpurun): settles what round 1 left unverified.
#   1. merge the side branch with the FFJORD forward wiring (written after round 1's GPU minutes were spent),
#   2. run its gated solve-level parity test next to the whole GPU suite,
#   3. print one default bench line.
# If step 2 is green, remove the RNDE_RUN_UNVERIFIED gate in tests/test_gpu_ffjord.py and keep the merge; otherwise
# `git merge --abort` / reset and debug csq_rhs inside fwd_kernel<..., FIELD = 1> with tools/sanitize_cases.py.
#
#   git merge --no-edit ffjord-stepper && python __graft_entry__.py          # here (CPU): merge + rebuild
#   gpurun --timeout 300 -- 'bash tools/next_round_first_gpu_call.sh'
set -euo pipefail
RNDE_RUN_UNVERIFIED=1 python -m pytest tests/test_gpu_ffjord.py -x -q
python -m pytest tests -x -q -m gpu
python bench.py --steps 20 --warmup 3 --no-cpu-baseline
