# A/B of the multi-GPU step tail (2 GPUs): slab-pipelined vs sequential all-reduce, NCCL stream priority / channel count.
mkdir -p gpurun_out/final
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/final/bench_2gpu_$name.json 2> gpurun_out/final/bench_2gpu_$name.err; python -c "import json;d=json.load(open('gpurun_out/final/bench_2gpu_$name.json'));print('$name',round(d['ms_per_step'],3),round(d['tail_ms']['value'],3),round(d['e2e']['ms_per_step'],3))"; }
run pipelined_hi UBN_BENCH_TAIL=pipelined
run pipelined_lo UBN_BENCH_TAIL=pipelined UBN_NCCL_HIGH_PRIORITY=0
run pipelined_hi_ch32 UBN_BENCH_TAIL=pipelined NCCL_MIN_NCHANNELS=32
run sequential_ch32 UBN_BENCH_TAIL=sequential NCCL_MIN_NCHANNELS=32
