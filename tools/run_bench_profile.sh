set -x
cd $GRAFT_REPO_ROOT
( time python bench.py > gpurun_out/r2_bench_n1c.json 2> gpurun_out/r2_bench_n1c.err ) 2> gpurun_out/r2_bench_n1c.time
tail -c 600 gpurun_out/r2_bench_n1c.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sketch_gemm_tf32_kernel -s 1 -c 1 -o gpurun_out/gemm32_r02 python tools/exp_gemm32.py --logn 22 --check 0 --iters 1 > gpurun_out/ncu_gemm32_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()"
