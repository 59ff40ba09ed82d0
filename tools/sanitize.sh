#!/bin/bash
# compute-sanitizer passes over a small host-driven stepping run (single engine and 2 shards in
# one process).  Writes one summary per tool under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for world in 1 2; do
    extra=""
    [ "$tool" = "synccheck" ] && extra="--num-cuda-barriers 16384"
    out=gpurun_out/san_${tool}_${world}.log
    timeout 240 $CS --tool $tool $extra --print-limit 20 python tests/tools/sanitize_target.py $world 2 > $out 2>&1
    echo "== $tool, $world shard(s): exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target|Error|error" $out | head -8
  done
done
