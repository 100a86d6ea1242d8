cd $GRAFT_REPO_ROOT
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/c18_bench2.json 2> gpurun_out/c18_bench2.err
tail -3 gpurun_out/c18_bench2.err; cut -c1-300 gpurun_out/c18_bench2.json
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1) > gpurun_out/c18_ref2.json 2> gpurun_out/c18_ref2.err
cut -c1-300 gpurun_out/c18_ref2.json; tail -2 gpurun_out/c18_ref2.err
