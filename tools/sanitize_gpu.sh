#!/bin/bash
# compute-sanitizer over tools/sanitize_scene.py (race detection). memcheck = out-of-bounds / misaligned accesses,
# racecheck = shared-memory hazards (block solver, TOI late rounds, class sort of the narrowphase), synccheck = barrier misuse,
# initcheck = reads of uninitialised device memory. Reports go to gpurun_out/sanitizer_<tool>.txt; copy them to profiles/.
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  steps=30; [ $tool = racecheck ] && steps=12
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_scene.py $steps > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize scene done|Error|hazard" gpurun_out/sanitizer_$tool.txt | sort | uniq -c | sort -rn | head -8
done
