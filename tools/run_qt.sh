echo "== reach, Shadow hand (BASELINE.json configs[0..1])"; python tools/gpu_reach_bench.py 2>&1 | tail -3
echo "== reach, Adroit hand (the reference's registered reach task)"; HAND=adroit_hand python tools/gpu_reach_bench.py 2>&1 | tail -3
echo "== IK (configs[3])"; python tools/gpu_ik_bench.py 2>&1 | tail -3
echo "== prop placement rounds"; python tools/gpu_place_bench.py 2>&1 | tail -3
