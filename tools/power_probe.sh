for v in v2 d2; do
  echo "== $v"
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active -lms 200 --format=csv,noheader > gpurun_out/pw_$v.csv &
  SMI=$!
  NBE_DBUF=0 NBE_LIB=$PWD/ab/$v.so timeout 120 python -u tools/run_one.py 224 150 2>/dev/null | egrep "per forward|total"
  kill $SMI
  sort gpurun_out/pw_$v.csv | uniq -c | sort -rn | head -8
done
