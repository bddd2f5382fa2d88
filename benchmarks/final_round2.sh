#!/bin/bash
# Round-2 single-GPU measurement set (run on the GPU box from the repo root); outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/bench1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/benchref.err
bash benchmarks/profile_step.sh
python benchmarks/fused_tiles.py > gpurun_out/r2_fused_tiles.jsonl 2> gpurun_out/tiles.err
python benchmarks/epoch_breakdown.py --epochs 5 --out gpurun_out/r2_epoch_breakdown.jsonl > /dev/null 2> gpurun_out/epoch.err
python benchmarks/epoch_breakdown.py --epochs 5 --per-minibatch --out gpurun_out/r2_epoch_breakdown_per_minibatch.jsonl > /dev/null 2>> gpurun_out/epoch.err
python benchmarks/ppo_epoch_multi.py --ctas 0 > gpurun_out/r2_ppo_epoch_1gpu.jsonl 2> gpurun_out/epoch1.err
rm -f gpurun_out/r2_ppo_readme_24env.jsonl
python -m ppo_car_b200.train_ppo --track big_track --n-envs 24 --n-epochs 200 --fused-rollout --fused-update \
    --log-json gpurun_out/r2_ppo_readme_24env.jsonl > gpurun_out/readme_run.log 2>&1
tail -2 gpurun_out/readme_run.log
