# GPU-box routine of a round's final build: tests, both bench arms, ncu launch list of the bench command, one full capture.
# usage: TAG=r02z bash tools/gpu_final.sh
set -x
TAG=${TAG:-r02z}
python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_err_$TAG.log | tee gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_err_$TAG.log | tee gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(tree8|collide|narrow|compact|dynamics)' -s 25 -c 6 -o gpurun_out/prof_$TAG -f python tools/profile_target.py > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
