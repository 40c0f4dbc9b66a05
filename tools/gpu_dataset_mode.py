#!/usr/bin/env python
"""Round-2 check (needs a B200; written after round 1's GPU minutes were spent, NOT yet run): `run_test.py --dataset` end to end
on a synthetic NOCS tree -- reader -> GpuCollate -> BPE tokenizer -> text tower -> network -> matching -> lifting -> PointDSC
-> gathered CSV -> `--score` (evaluator incl. VSD) -- with seeded random weights and the synthetic BPE vocabulary of
tests/golden.  The architecture fixes the number of templated prompts at 80 (fusion `prompt_channel`), so the tree's
templates.json is rewritten with 80 templates.

    gpurun -- 'python tools/gpu_dataset_mode.py'                                   (1 GPU)
    gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_dataset_mode.py'
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oryon_b200 import synth  # noqa: E402
import run_test  # noqa: E402


def main():
    """`--repeat R`: throughput mode -- the 6-pair list of the synthetic tree is repeated R times (the same PNG files are decoded
    again for every sample, as a real split's distinct files would be), batches of 32, no scoring: the whole-loop pairs/s with
    decoding on the loader's worker threads, i.e. what a from-disk run delivers on this host."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    repeat = int(sys.argv[sys.argv.index("--repeat") + 1]) if "--repeat" in sys.argv else 1
    workers = sys.argv[sys.argv.index("--workers") + 1] if "--workers" in sys.argv else "8"
    d = os.environ.get("ORYON_TREE") or os.path.join(tempfile.gettempdir(), f"oryon_nocs_tree_r{repeat}")
    done = os.path.join(d, "tree.done")
    if rank == 0 and not os.path.exists(done):
        info = synth.write_nocs_tree(d, 0, hw=(480, 640))
        with open(os.path.join(info["base"], "templates.json"), "w") as f:
            json.dump([f"a photo number {i} of a {{}}." for i in range(80)], f)
        if repeat > 1:
            lst = os.path.join(info["base"], "fixed_split", info["split"], "instance_list.txt")
            lines = open(lst).readlines()
            open(lst, "w").writelines(lines * repeat)
        open(done, "w").write("ok")
    while not os.path.exists(done):      # the other ranks wait for rank 0's tree
        import time
        time.sleep(0.2)
    out = os.path.join(ROOT, "gpurun_out", f"dataset_mode_pred_r{repeat}_n{world}.csv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    args = ["--dataset", "nocs", "--root", d, "--split", "cross_scene_test", "--obj", "all", "--mask", "oracle",
            "--batch", "4" if repeat == 1 else "32", "--workers", workers,
            "--bpe", os.path.join(ROOT, "tests", "golden", "bpe_synth_vocab.txt.gz"), "--out", out]
    run_test.main(args + (["--score"] if repeat == 1 else []))
    if rank == 0 and repeat == 1:
        print(open(out).read(), file=sys.stderr)
        print(open(os.path.splitext(out)[0] + ".json").read()[:600], file=sys.stderr)


if __name__ == "__main__":
    main()
