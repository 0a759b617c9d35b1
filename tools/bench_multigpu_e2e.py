"""End-to-end OCR of one page set sharded over N GPUs (one process per GPU, page i -> rank i mod N, host-side gather):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \\
        tools/bench_multigpu_e2e.py [pages]

Rank 0 prints one JSON line and checks the gathered result against its own single-GPU run of all pages."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from onnxocr_b200 import multigpu, pipeline, synth_pages  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
host_group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    host_group = dist.new_group(backend="gloo")     # the result gather is host-side work: keep it off the GPUs
ocr = pipeline.ONNXPaddleOcr(use_angle_cls=True, precision="bf16", gpu_id=local)
pages = synth_pages.e2e_pages_c4(n, seed=2)
mine = multigpu.shard_indices(n, rank, world)


def run_shard():
    res = ocr.ocr_batch([pages[i] for i in mine])
    return list(zip(mine, res))


for _ in range(2):
    run_shard()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
t_run = t_gather = 0.0
for _ in range(reps):
    ta = time.perf_counter()
    local_res = run_shard()
    tb = time.perf_counter()
    gathered = multigpu.gather_by_page(local_res, n, group=host_group)
    t_run += tb - ta
    t_gather += time.perf_counter() - tb
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / reps
if rank == 0:
    single = ocr.ocr_batch(pages)
    same = sum(1 for a, b in zip(gathered, single) if [x[1][0] for x in a] == [x[1][0] for x in b])
    lines = sum(len(r) for r in gathered)
    print(json.dumps({"config": "C4 end-to-end det+cls+rec, %d mixed-size pages sharded page i -> rank i mod N" % n, "n_gpus": world,
                      "pages_per_s": n / dt, "rank0_ms_per_rep": {"ocr_batch_of_shard": 1e3 * t_run / reps, "gather": 1e3 * t_gather / reps}, "lines_per_s": lines / dt, "pages_identical_to_single_gpu": "%d/%d" % (same, n)}))
if world > 1:
    dist.destroy_process_group()
