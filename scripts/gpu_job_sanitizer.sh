#!/bin/bash
# compute-sanitizer over the resident-path tests (memcheck, synccheck, racecheck)
T=$1
SEL="tests/test_resident.py::test_resident_random_structures tests/test_resident.py::test_resident_stages tests/test_resident.py::test_single_cta_bucketing_equals_radix_sort_path tests/test_resident.py::test_resident_hands_over_to_tile_pipeline"
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -x -q > gpurun_out/${T}_sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_$tool.txt | tail -3
done
