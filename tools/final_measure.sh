set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --kernels > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 600 gpurun_out/bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>/dev/null; tail -c 400 gpurun_out/bench_ref.json
timeout 200 python tools/bench_density.py 512 > gpurun_out/density_bench.json 2>&1; cat gpurun_out/density_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/per_layer.csv python tools/run_one.py 224 1 > /dev/null 2>&1; wc -l gpurun_out/per_layer.csv
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2 -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-alt --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/launches_bench.csv
