"""Parity / timing probe of the tensor-core op (one process per case so that a device trap cannot take the rest down).

  python tools/r2_probe.py parity B Cq C H W fp32|bf16 [samples]     -> one JSON line of max errors vs the fp64 oracle
  python tools/r2_probe.py time   B Cq C H W fp32|bf16 [iters]       -> fwd / bwd ms (L2 flushed), roofline fractions
Knobs come from the environment (CCA_B200_PDL, CCA_B200_ZERO_AHEAD, CCA_B200_DELTA)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ccnet_b200 import cca_backward, cca_forward


def main():
    mode = sys.argv[1]
    B, Cq, C, H, W = (int(a) for a in sys.argv[2:7])
    dt = {"fp32": torch.float32, "bf16": torch.bfloat16}[sys.argv[7]]
    extra = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 1000 + H * 7 + W)
    sc = 0.6
    q = (torch.randn(B, Cq, H, W, generator=g) * sc).to(dt)
    k = (torch.randn(B, Cq, H, W, generator=g) * sc).to(dt)
    v = torch.randn(B, C, H, W, generator=g).to(dt)
    do = torch.randn(B, C, H, W, generator=g).to(dt)
    cl = torch.channels_last
    qd, kd, vd, dd = (t.to(dev).contiguous(memory_format=cl) for t in (q, k, v, do))
    knobs = {n: os.environ.get(n) for n in ("CCA_B200_PDL", "CCA_B200_ZERO_AHEAD", "CCA_B200_DELTA", "CCA_B200_LAG", "CCA_B200_L2HINT") if os.environ.get(n)}
    rec = {"mode": mode, "shape": [B, Cq, C, H, W], "dtype": sys.argv[7], "knobs": knobs}
    if mode == "parity":
        from oracle import cca_oracle as O
        out, lse = cca_forward(qd, kd, vd, impl="tc")
        dq, dk, dv = cca_backward(dd, qd, kd, vd, out, lse, impl="tc")
        out2, lse2 = cca_forward(qd, kd, vd, impl="tc")
        g2 = cca_backward(dd, qd, kd, vd, out2, lse2, impl="tc")
        rec["rerun_bit_identical"] = bool(torch.equal(out, out2) and torch.equal(lse, lse2)
                                          and all(torch.equal(a, b) for a, b in zip((dq, dk, dv), g2)))
        torch.cuda.synchronize()
        samples = sorted(set([0, B - 1] + ([B // 2] if B > 2 else [])))[: max(1, extra) if extra else 3]
        errs = {"out": 0.0, "lse": 0.0, "dq": 0.0, "dk": 0.0, "dv": 0.0}
        rel = dict(errs)
        for b in samples:
            sl = slice(b, b + 1)
            ro, rl = O.cca_forward(q[sl].double(), k[sl].double(), v[sl].double())
            rq, rk, rv = O.cca_backward(do[sl].double(), q[sl].double(), k[sl].double(), v[sl].double())
            for name, got, ref in (("out", out, ro), ("lse", lse, rl), ("dq", dq, rq), ("dk", dk, rk), ("dv", dv, rv)):
                e = (got[sl].cpu().double() - ref).abs().max().item()
                errs[name] = max(errs[name], e)
                rel[name] = max(rel[name], e / max(1.0, ref.abs().max().item()))
        rec.update({"samples": samples, "max_abs_err": errs, "max_err_rel_to_max_ref": rel,
                    "finite": bool(torch.isfinite(out).all() and torch.isfinite(dv).all())})
    else:
        iters = extra or 10
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        out, lse = cca_forward(qd, kd, vd, impl="tc")

        def op_time(fn):
            ts = []
            for _ in range(3):
                fn()
            for _ in range(iters):
                flush.zero_()
                torch.cuda._sleep(300000)      # GPU spin while the host enqueues the op's launches (no host latency in the events)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            return sum(ts) / len(ts), ts[0]

        f, fmin = op_time(lambda: cca_forward(qd, kd, vd, impl="tc"))
        b, bmin = op_time(lambda: cca_backward(dd, qd, kd, vd, out, lse, impl="tc"))
        es = 4 if dt == torch.float32 else 2
        N = B * H * W
        bf, bb = es * N * (2 * Cq + 2 * C), es * N * (4 * Cq + 4 * C)
        peak = 6484.3
        rec.update({"fwd_ms": f, "fwd_ms_min": fmin, "bwd_ms": b, "bwd_ms_min": bmin,
                    "fwd_frac": bf / f / 1e6 / peak, "bwd_frac": bb / b / 1e6 / peak})
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
