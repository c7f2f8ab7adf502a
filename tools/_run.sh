set -x
python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01c.json 2> gpurun_out/bench_ref_r01c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_r01c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_schur_pairs|k_lin_points_pipe|k_backsub_pipe|k_lin_cams|k_panel_step|k_backward_flow|k_S_finalize|k_vinv" -c 12 -o gpurun_out/prof_r01c -f python tools/prof_run.py full 1 > gpurun_out/ncu_full_r01c.log 2>&1
tail -2 gpurun_out/ncu_full_r01c.log
