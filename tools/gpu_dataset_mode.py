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
    rank = int(os.environ.get("RANK", "0"))
    d = os.environ.get("ORYON_TREE") or os.path.join(tempfile.gettempdir(), "oryon_nocs_tree")
    if rank == 0 and not os.path.exists(os.path.join(d, "nocs", "templates.json")):
        info = synth.write_nocs_tree(d, 0, hw=(480, 640))
        with open(os.path.join(info["base"], "templates.json"), "w") as f:
            json.dump([f"a photo number {i} of a {{}}." for i in range(80)], f)
    out = os.path.join(ROOT, "gpurun_out", "dataset_mode_pred.csv")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    run_test.main(["--dataset", "nocs", "--root", d, "--split", "cross_scene_test", "--obj", "all", "--mask", "oracle", "--batch", "4",
                   "--bpe", os.path.join(ROOT, "tests", "golden", "bpe_synth_vocab.txt.gz"), "--out", out, "--score"])
    if rank == 0:
        print(open(out).read(), file=sys.stderr)
        print(open(os.path.splitext(out)[0] + ".json").read()[:600], file=sys.stderr)


if __name__ == "__main__":
    main()
