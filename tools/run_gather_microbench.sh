#!/bin/bash
# Runs tools/build/gather_microbench over the variants of interest, one process per variant under `timeout`
# (a faulting variant cannot take the others with it).  Output: one JSON line per variant in $1.
OUT=${1:-gpurun_out/gather_microbench.jsonl}
B=tools/build/gather_microbench
: > "$OUT"
run() { timeout 90 $B "$@" >> "$OUT" 2>> "${OUT%.jsonl}.err" || echo "{\"args\": \"$*\", \"failed\": $?}" >> "$OUT"; }
for c in 4 6 8; do run lsu $c; done
for s in 2 4 8; do run lds 4 $s; done
for c in 2 4 6; do for s in 3 4 6; do run bulk $c $s; done; done
run bulk 3 8
run gather4 4 4 0 0 1
run gather4 4 4 0 0 4
run gather4 6 4 0 0 1
run gather4 3 8 0 0 1
for f in 0.5 0.6 0.7 0.8; do run mix 4 4 $f 4; run mix 5 4 $f 5; done
run mix 4 6 0.6 4
run mix 6 3 0.7 5
cat "$OUT"
