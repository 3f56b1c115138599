#!/bin/bash
# compute-sanitizer passes over the hand-written kernels (run on the GPU box):
#   tools/sanitize.sh [outdir]      -> <outdir>/sanitizer_<tool>_<case>.log + sanitizer_summary.txt
# memcheck: out-of-bounds / misaligned global + shared accesses (incl. the TMA tile copies of K4c and the
# peer ring); racecheck: shared-memory hazards (block reductions, K4c stages, KB row exchange);
# synccheck: barrier misuse; initcheck: reads of uninitialised device memory.
out=${1:-gpurun_out}
mkdir -p "$out"
cs=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
: > "$out/sanitizer_summary.txt"
rc_all=0
for tool in memcheck racecheck synccheck initcheck; do
  for c in smoke team recompute; do
    log="$out/sanitizer_${tool}_${c}.log"
    extra=""
    [ "$tool" = memcheck ] && extra="--leak-check no"
    [ "$tool" = initcheck ] && extra="--track-unused-memory no"
    timeout 900 "$cs" --tool "$tool" $extra --error-exitcode 97 --print-limit 20 \
        python tools/sanitize_case.py "$c" > "$log" 2>&1
    rc=$?
    tail_line=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    echo "$tool $c rc=$rc :: $tail_line" | tee -a "$out/sanitizer_summary.txt"
    [ $rc -ne 0 ] && rc_all=1
  done
done
exit $rc_all
