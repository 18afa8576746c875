mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 2> gpurun_out/bench_n2.err | tee gpurun_out/bench_n2.json | cut -c1-700
tail -3 gpurun_out/bench_n2.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-900
