#!/bin/bash
# Development aid (on a GPU box): compute-sanitizer over a subset of the GPU parity tests.
#   scripts/sanitize.sh memcheck | racecheck
set -e
tool=${1:-memcheck}
sel='golden or modeac_matches or ragged or adversarial or span_split or try_masks or converter'
[ "$tool" = racecheck ] && sel='uc8_fix1 or uc8_modeac or sc16q11 or adversarial'
exec compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel"
