#!/bin/bash
# Sample SM clock / board power / throttle reasons every 200 ms while tools/run_one.py runs 150
# forwards of the 224^3 Style+vel subbox, once per library given on the command line
# (default: the in-tree build).  Usage on the GPU box:  bash tools/power_probe.sh [lib.so ...]
# The "zero data" comparison in profiles/r1_power/ used a debug build whose epilogue stores were
# redirected into a 512 KB window (so every layer after the first reads zeros); see DESIGN.md section 6.
libs=("$@"); [ ${#libs[@]} -eq 0 ] && libs=("$PWD/jax_nbody_emulator_with_dj_b200/libnbe_b200.so")
mkdir -p gpurun_out
for lib in "${libs[@]}"; do
  tag=$(basename "$lib" .so)
  echo "== $tag"
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active \
    -lms 200 --format=csv,noheader > "gpurun_out/pw_$tag.csv" &
  SMI=$!
  NBE_LIB="$lib" timeout 120 python -u tools/run_one.py 224 150 2>/dev/null | egrep "per forward|total"
  kill $SMI
  awk -F, '$3+0 > 900' "gpurun_out/pw_$tag.csv" | sort | uniq -c | sort -rn | head -5
done
