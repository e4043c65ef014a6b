"""CLI around oracle/ref_harness.py: run the UNMODIFIED reference sampler once and save the final latent.
TEST INFRASTRUCTURE (used by tests/test_final_latent_gpu.py and scripts/ref_gpu_compare.py through a subprocess, because
the test process has this repo's drop-in `ldm` imported and the reference's `ldm` needs a clean interpreter).

    python oracle/ref_run.py --config sd14_box_text --S 50 --alpha 0.3,0,0.7 --device cuda:0 --autocast bf16 --out lat.pt

Weights: gligen_b200.spec.synthetic_state_dict(cfg, seed 0); inputs: gligen_b200.synth.make_inputs(cfg, B, max_objs, seed)
- the same bits the golden fixtures (oracle/gen_golden.py) and the engine tests use.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sd14_box_text")
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--max-objs", type=int, default=30)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--kind", default="plms", choices=["plms", "ddim"])
    ap.add_argument("--S", type=int, default=50)
    ap.add_argument("--alpha", default="1,0,0")
    ap.add_argument("--guidance", type=float, default=7.5)
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--autocast", default="none", choices=["none", "bf16", "fp16"], help="comma list allowed: none,bf16")
    ap.add_argument("--autocasts", default="", help="comma list of precisions to run in ONE process (model built once)")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    from oracle import ref_harness as RH
    RH.mount()
    from gligen_b200 import synth
    from gligen_b200.spec import NAMED_CONFIGS, synthetic_state_dict
    cfg = NAMED_CONFIGS[args.config]
    sd = synthetic_state_dict(cfg, seed=0)
    inp = synth.make_inputs(cfg, args.B, args.max_objs, seed=args.seed)
    atype = [float(v) for v in args.alpha.split(",")]
    precisions = [p for p in (args.autocasts.split(",") if args.autocasts else [args.autocast]) if p]
    dt_map = {"none": None, "bf16": torch.bfloat16, "fp16": torch.float16}
    out = {"config": args.config, "B": args.B, "max_objs": args.max_objs, "S": args.S, "alpha_type": atype, "kind": args.kind,
           "device": args.device, "latents": {}, "seconds": {}}
    for prec in precisions:
        lat, dt = RH.run_reference_sampler(cfg, sd, inp, args.kind, args.S, atype, args.guidance, args.device, dt_map[prec])
        out["latents"][prec] = lat
        out["seconds"][prec] = dt
    torch.save(out, args.out)
    print(f"wrote {args.out}: " + ", ".join(f"{k} {v:.1f}s" for k, v in out["seconds"].items()))


if __name__ == "__main__":
    main()
