"""Shared helpers of the parity tests: fixtures, module/oracle runners, the error metric.

Metric: max|got - ref| / max|ref| per tensor.  Tolerance for the FP32 path is
1e-4 (BASELINE.json north_star).  A conv bias that feeds a BatchNorm has an exactly-zero true
gradient, so bias gradients are measured against the magnitude of the sibling weight gradient.
"""
import os

import numpy as np
import torch

import cases                      # tests/golden/cases.py
from oracle import ffc_ref as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-4                        # north_star: FP32 path within 1e-4 max relative error


def load_fixture(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


def relerr(got, ref, floor=0.0):
    got = torch.as_tensor(got).double().cpu()
    ref = torch.as_tensor(ref).double().cpu()
    if ref.numel() == 0:
        return 0.0
    assert got.shape == ref.shape, (got.shape, ref.shape)
    den = max(ref.abs().max().item(), floor, 1e-30)
    return (got - ref).abs().max().item() / den


def fixture_state(fx, dtype=torch.float32):
    sd = {}
    for k, v in fx.items():
        if k.startswith("sd/"):
            t = torch.from_numpy(v.copy())
            sd[k[3:]] = t.to(dtype) if t.is_floating_point() else t
    return sd


def fixture_inputs(fx, dtype=torch.float32):
    xs = []
    i = 0
    while f"in{i}" in fx:
        xs.append(torch.from_numpy(fx[f"in{i}"].copy()).to(dtype))
        i += 1
    return xs


def _collect(outs, xs, named_params, named_buffers):
    res = {}
    for i, o in enumerate(outs):
        res[f"out{i}"] = o.detach()
    for i, x in enumerate(xs):
        res[f"din{i}"] = x.grad
    for k, p in named_params:
        if p.grad is not None:
            res["grad/" + k] = p.grad
    for k, b in named_buffers:
        res["post/" + k] = b.detach()
    return res


def run_module(mod, fx, device="cpu"):
    """Loads the fixture's state into ``mod`` (a fastfourierconvolution_b200 module), runs forward and
    backward with the fixture's inputs / cotangents on ``device``."""
    mod.load_state_dict(fixture_state(fx), strict=True)
    mod.to(device)
    mod.train(bool(fx["train"]))
    xs = [x.to(device).requires_grad_(True) for x in fixture_inputs(fx)]
    out = mod(xs[0] if len(xs) == 1 else tuple(xs))
    outs = [o for o in (out if isinstance(out, tuple) else (out,)) if torch.is_tensor(o)]
    loss = sum((o * torch.from_numpy(fx[f"cot{i}"]).to(device)).sum() for i, o in enumerate(outs))
    loss.backward()
    return _collect(outs, xs, mod.named_parameters(), mod.named_buffers())


def run_oracle(name, fx, dtype=torch.float64):
    """Runs oracle/ffc_ref.py on the fixture (CPU)."""
    oracle = cases.CASES[name][3](R)
    P = {}
    for k, v in fixture_state(fx, dtype).items():
        P[k] = v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(
            ("running_mean", "running_var", "weight_u", "weight_v")) else v.clone()
    xs = [x.requires_grad_(True) for x in fixture_inputs(fx, dtype)]
    out = oracle(P, xs, bool(fx["train"]))
    outs = [o for o in (out if isinstance(out, tuple) else (out,)) if torch.is_tensor(o)]
    loss = sum((o * torch.from_numpy(fx[f"cot{i}"]).to(dtype)).sum() for i, o in enumerate(outs))
    loss.backward()
    params = [(k, v) for k, v in P.items() if v.requires_grad]
    buffers = [(k, v) for k, v in P.items() if not v.requires_grad]
    return _collect(outs, xs, params, buffers)


def compare(got, ref, tol=TOL, what=""):
    """ref: dict of reference arrays (fixture or oracle result).  Returns {key: err}; asserts tol."""
    errs = {}
    for k, r in ref.items():
        if not (k.startswith(("out", "din", "grad/", "post/"))):
            continue
        r = torch.as_tensor(r)
        if not r.is_floating_point():
            assert k in got and torch.equal(torch.as_tensor(got[k]).cpu(), r), f"{what}: {k} differs"
            continue
        assert k in got, f"{what}: {k} missing (have {sorted(got)})"
        floor = 0.0
        if k.startswith("grad/") and k.endswith("bias"):
            sib = k[:-4] + "weight"
            sib = sib if sib in ref else k[:-4] + "weight_orig"
            if sib in ref:
                floor = float(np.abs(np.asarray(torch.as_tensor(ref[sib]).double())).max())
        errs[k] = relerr(got[k], r, floor)
    for k in got:
        if k.startswith("grad/"):
            assert k in ref, f"{what}: unexpected gradient for {k}"
    bad = {k: v for k, v in errs.items() if not v <= tol}
    assert not bad, f"{what}: above tolerance {tol}: {bad}"
    return errs


# ---------------------------------------------------------------------------------------------
# mask-flip-aware comparison (SURVEY.md section 8(c) caveat 1)
# ---------------------------------------------------------------------------------------------
def _grad_floor(k, ref):
    """Bias gradients in front of a BatchNorm are exactly zero: measure them against the sibling weight's gradient."""
    if k.startswith("grad/") and k.endswith("bias"):
        for sib in (k[:-4] + "weight", k[:-4] + "weight_orig"):
            if sib in ref:
                return float(torch.as_tensor(ref[sib]).double().abs().max())
    return 0.0


def flip_aware_compare(got, oracle_run, ref=None, tol=TOL, out_tol=None, margins=(4e-6, 4e-5), max_flips=40, what=""):
    """Holds EVERY output, buffer and gradient of a deep network to ``tol`` in the max norm (max|d| / max|ref|), while
    accounting for ReLU / LeakyReLU elements that the two FP32-accurate evaluations put on different sides of the kink.

    ``oracle_run(overrides) -> (result dict, ActTape)`` runs the float64 oracle with the given mask overrides.  ``ref`` is
    what ``got`` is compared with: the oracle's own unperturbed result (default) or a golden fixture of the reference.

    1. Plain comparison.  Nothing above ``tol``: done, zero flips.
    2. Otherwise the elements whose float64 pre-activation lies within ``margin`` (relative to the site's max) of the
       kink are the only ones a rounding difference can flip.  Each candidate is flipped alone in an oracle replay; the
       change of all gradients it causes, D_e, is exact and the flips superpose (masks are piecewise constant).
    3. Least squares for the flip indicators c_e over all gradient entries, rounded to integers (0 / 1 against the
       oracle, -1 / 0 / 1 against a fixture, whose own FP32 run may have flipped too).  The gradients corrected by the
       detected flips must then meet ``tol`` everywhere.  Forward values are never corrected.
    Returns (errs, n_flips)."""
    res0, tape = oracle_run({})
    ref = res0 if ref is None else ref
    out_tol = tol if out_tol is None else out_tol
    keys = [k for k in ref if k.startswith(("out", "din", "grad/", "post/")) and k in res0
            and torch.as_tensor(ref[k]).is_floating_point() and torch.as_tensor(ref[k]).numel() > 0]
    for k in keys:
        assert k in got and got[k] is not None, f"{what}: {k} missing in the product's results"
    is_grad = lambda k: k.startswith(("din", "grad/"))
    scale = {k: max(float(torch.as_tensor(ref[k]).double().abs().max()), _grad_floor(k, ref), 1e-30) for k in keys}
    diff = {k: (torch.as_tensor(got[k]).detach().double().cpu() - torch.as_tensor(ref[k]).double()).reshape(-1) / scale[k] for k in keys}
    errs = {k: float(d.abs().max()) if d.numel() else 0.0 for k, d in diff.items()}
    lim = lambda k: tol if is_grad(k) else out_tol
    bad = {k: v for k, v in errs.items() if not v <= lim(k)}
    if not bad:
        return errs, 0
    assert all(is_grad(k) for k in bad), f"{what}: forward values above tolerance: { {k: v for k, v in bad.items() if not is_grad(k)} }"
    gkeys = [k for k in keys if is_grad(k)]
    last = None
    for margin in margins:
        cand = []
        for i, pre in enumerate(tape.pre):
            a = pre.abs().reshape(-1)
            if a.numel() == 0:
                continue
            thr = margin * float(a.max())
            for j in torch.nonzero(a < thr).reshape(-1).tolist():
                cand.append((float(a[j]) / max(float(a.max()), 1e-300), i, j))
        cand.sort()
        cand = cand[:max_flips]
        if not cand:
            continue
        cols = []
        for _, i, j in cand:
            mask = (tape.pre[i] > 0).clone()
            mask.view(-1)[j] = ~mask.view(-1)[j]
            res_e, _ = oracle_run({i: mask})
            cols.append({k: ((res_e[k].double() - res0[k].double()).reshape(-1) / scale[k]).float() for k in gkeys})
        K = len(cols)
        AtA = torch.zeros(K, K, dtype=torch.float64)
        Atb = torch.zeros(K, dtype=torch.float64)
        for k in gkeys:
            A = torch.stack([c[k] for c in cols], 1).double()
            AtA += A.T @ A
            Atb += A.T @ diff[k]
        c = torch.linalg.lstsq(AtA + 1e-18 * torch.eye(K, dtype=torch.float64), Atb.unsqueeze(1)).solution.reshape(-1)
        cr = c.round().clamp(-1, 1)
        fixed = {}
        for k in gkeys:
            A = torch.stack([cc[k] for cc in cols], 1).double()
            fixed[k] = float((diff[k] - A @ cr).abs().max()) if diff[k].numel() else 0.0
        last = (fixed, int((cr != 0).sum()), [(round(float(x), 3)) for x in c.tolist()], margin, K)
        if all(v <= tol for v in fixed.values()):
            errs.update(fixed)
            return errs, last[1]
    assert False, (f"{what}: gradients above tolerance {tol} that no activation-mask flip explains: {bad}; "
                   f"after flip fitting: {None if last is None else {k: v for k, v in last[0].items() if v > tol}} "
                   f"(flips {None if last is None else last[1:]})")
