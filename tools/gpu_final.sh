set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_err.log | tee gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err.log | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1s.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(tree8|collide|narrow|compact|dynamics)' -s 25 -c 5 -o gpurun_out/prof_r1s -f python tools/profile_target.py > gpurun_out/ncu_r1s.log 2>&1
tail -2 gpurun_out/ncu_r1s.log
python tools/gpu_place_bench.py 2>&1 | tail -2 | tee gpurun_out/place_bench.log
