cd $GRAFT_REPO_ROOT
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 3 --warmup 3) > gpurun_out/c27_bench4.json 2> gpurun_out/c27_bench4.err
tail -3 gpurun_out/c27_bench4.err | cut -c1-300; cut -c1-300 gpurun_out/c27_bench4.json
