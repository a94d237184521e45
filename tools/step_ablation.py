"""Where the step time goes beyond the sum of the main-stream kernels: time the graphed training step (configs[1], batch 64)
with one piece of side-stream work removed or moved at a time.  Measurement tool only (not a product path).

  python tools/step_ablation.py [--steps 30]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def build(variant):
    from fairfedmed_b200 import ops
    from fairfedmed_b200.registry import build_trainer
    import fairfedmed_b200.trainer  # noqa: F401
    from fairfedmed_b200 import clip_model
    c = bench.CONFIGS[2]
    cfg = bench.make_cfg(1, bench.BATCH, "Sinkhorn", c)
    cfg.SEED = 1
    ops.PARAMS_ON_SIDE_STREAM = variant != "adapter_grads_on_main"
    clip_model.Transformer.hoist_adapter_prep = variant != "no_hoisted_prep"
    tr = build_trainer(cfg)
    tr.sync_metrics = False
    tr.step_auc = False
    tr.model.check_nan = False
    tr.batch_idx, tr.num_batches = 0, 10 ** 9
    tr.model.overlap_text = variant != "text_on_main"
    if variant == "no_text_tower":
        with torch.no_grad():
            txt = tr.model.text_encoder(tr.model.prompt_learner(), tr.model.prompt_learner.eot_index).detach()

        class _Cached(torch.nn.Module):
            def forward(self, prompts, eot):
                return txt + 0.0 * prompts.sum()
        tr.model.text_encoder = _Cached()
    with torch.no_grad():
        g = torch.Generator().manual_seed(7)
        for n_, p_ in tr.model.named_parameters():
            if "lora_A" in n_:
                p_.copy_((0.02 * torch.randn(p_.shape, generator=g)).to("cuda:0"))
    return tr, c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--variants", default="baseline,no_text_tower,text_on_main,adapter_grads_on_main,no_hoisted_prep")
    args = ap.parse_args()
    from fairfedmed_b200.config import ATTRIBUTE_GROUPS
    for variant in args.variants.split(","):
        tr, c = build(variant)
        gen = torch.Generator().manual_seed(100)
        groups = [len(ATTRIBUTE_GROUPS[c["dataset"]][a]) for a in c["attributes"]]
        pool = [{k: v.to("cuda:0") for k, v in bench.synthetic_batch(gen, bench.BATCH, c, groups).items()} for _ in range(4)]
        for i in range(2):
            tr.forward_backward(pool[i])
        tr.capture_step_graph(pool[0])
        for i in range(5):
            tr.forward_backward_graphed(pool[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            tr.forward_backward_graphed(pool[i % 4])
        e1.record()
        torch.cuda.synchronize()
        print(f"{variant:24s} {e0.elapsed_time(e1) / args.steps:7.3f} ms/step", flush=True)
        del tr
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
