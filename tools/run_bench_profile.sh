set -x
cd $GRAFT_REPO_ROOT
( time python bench.py > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err ) 2> gpurun_out/r2_bench_n1b.time
tail -c 600 gpurun_out/r2_bench_n1b.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_b2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sketch_gemm_kernel -s 2 -c 1 -o gpurun_out/gemm_r02 python tools/bench_gemm.py --logn 22 --iters 2 > gpurun_out/ncu_gemm_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep
