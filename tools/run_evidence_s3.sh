set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_final.log 2>&1; tail -2 gpurun_out/gputest_final.log
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_final_n1.err
python tools/profile_factor.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:gs_grid -c 1 -o gpurun_out/gs_grid_r02 python tools/profile_factor.py > gpurun_out/ncu_gs_grid.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
ls -la gpurun_out/ | tail -12
