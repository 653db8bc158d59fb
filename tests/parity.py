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


def _fit_flips(AtA, Atb, lo, hi):
    """Integer flip indicators c (lo <= c_e <= hi) that minimise |d - A c|_2, by greedy coordinate moves of +-1 in the Gram
    representation (AtA = A^T A, Atb = A^T d).  Robust to (nearly) collinear candidates -- e.g. the four copies a nearest-x2
    upsampling makes of one pre-activation -- where rounding an unconstrained least-squares solution is not."""
    K = Atb.numel()
    c = torch.zeros(K, dtype=torch.float64)
    g = Atb.clone()                                 # A^T (d - A c)
    diag = torch.diagonal(AtA)
    for _ in range(4 * K + 8):
        # moving c_e by s in {+1, -1} changes |r|^2 by diag_e - 2 s g_e
        gain_p = torch.where(c < hi, 2 * g - diag, torch.full_like(g, -1.0))
        gain_m = torch.where(c > lo, -2 * g - diag, torch.full_like(g, -1.0))
        bp, bm = int(torch.argmax(gain_p)), int(torch.argmax(gain_m))
        if gain_p[bp] <= 0 and gain_m[bm] <= 0:
            break
        if gain_p[bp] >= gain_m[bm]:
            c[bp] += 1; g -= AtA[:, bp]
        else:
            c[bm] -= 1; g += AtA[:, bm]
    return c


def flip_aware_errors(got, oracle_run, ref=None, margins=(2e-6, 1e-5, 5e-5, 2e-4), max_flips=192, good=TOL, cache=None):
    """Max-norm errors (max|d| / max|ref| per tensor) of ``got`` after accounting for activation-mask flips; no assertion.
    Returns (errs, n_flips, plain_errs).  See flip_aware_compare.  ``cache`` (a dict) shares the oracle's base run and the
    per-candidate gradient responses between several calls on the same oracle_run."""
    cache = {} if cache is None else cache
    if "base" not in cache:
        cache["base"] = oracle_run({})
    res0, tape = cache["base"]
    against_fixture = ref is not None
    ref = res0 if ref is None else ref
    keys = [k for k in ref if k.startswith(("out", "din", "grad/", "post/")) and k in res0
            and torch.as_tensor(ref[k]).is_floating_point() and torch.as_tensor(ref[k]).numel() > 0]
    for k in keys:
        assert k in got and got[k] is not None, f"{k} missing in the results under test"
    is_grad = lambda k: k.startswith(("din", "grad/"))
    scale = {k: max(float(torch.as_tensor(ref[k]).double().abs().max()), _grad_floor(k, ref), 1e-30) for k in keys}
    diff = {k: (torch.as_tensor(got[k]).detach().double().cpu() - torch.as_tensor(ref[k]).double()).reshape(-1) / scale[k] for k in keys}
    plain = {k: float(d.abs().max()) if d.numel() else 0.0 for k, d in diff.items()}
    gkeys = [k for k in keys if is_grad(k)]
    if all(plain[k] <= good for k in gkeys):
        return dict(plain), 0, plain
    if "ranked" not in cache:
        ranked = []
        for i, pre in enumerate(tape.pre):
            a = pre.abs().reshape(-1)
            if a.numel() == 0:
                continue
            rms = max(float(a.double().pow(2).mean().sqrt()), 1e-300)       # RMS of the site
            for j in torch.nonzero(a < margins[-1] * rms).reshape(-1).tolist():
                ranked.append((float(a[j]) / rms, i, j))
        ranked.sort()
        cache["ranked"], cache["cols"] = ranked[:max_flips], {}
    ranked, colcache = cache["ranked"], cache["cols"]
    best = (dict(plain), 0)
    for margin in margins:
        cand = [r for r in ranked if r[0] < margin]
        if not cand:
            continue
        cols = []
        for _, i, j in cand:
            if (i, j) not in colcache:
                mask = (tape.pre[i] > 0).clone()
                mask.view(-1)[j] = ~mask.view(-1)[j]
                res_e, _ = oracle_run({i: mask})
                # response of every gradient to this one flip, in units of the ORACLE's own maxima (rescaled per call below)
                colcache[(i, j)] = {k: (res_e[k].double() - res0[k].double()).reshape(-1).float() for k in res0 if is_grad(k)}
            cols.append(colcache[(i, j)])
        K = len(cols)
        AtA = torch.zeros(K, K, dtype=torch.float64)
        Atb = torch.zeros(K, dtype=torch.float64)
        mats = {}
        for k in gkeys:
            mats[k] = torch.stack([c[k] for c in cols], 1).double() / scale[k]
            AtA += mats[k].T @ mats[k]
            Atb += mats[k].T @ diff[k]
        cr = _fit_flips(AtA, Atb, -1.0 if against_fixture else 0.0, 1.0)
        fixed = dict(plain)
        for k in gkeys:
            fixed[k] = float((diff[k] - mats[k] @ cr).abs().max()) if diff[k].numel() else 0.0
        if max(fixed[k] for k in gkeys) < max(best[0][k] for k in gkeys):
            best = (fixed, int((cr != 0).sum()))
        if all(fixed[k] <= good for k in gkeys):
            break
    return best[0], best[1], plain


def flip_aware_compare(got, oracle_run, ref=None, tol=TOL, out_tol=None, margins=(2e-6, 1e-5, 5e-5, 2e-4), max_flips=192, what="",
                       noise_floor=None, cache=None):
    """Holds EVERY output, buffer and gradient of a deep network to ``tol`` in the max norm (max|d| / max|ref|), while
    accounting for ReLU / LeakyReLU elements that the two FP32-accurate evaluations put on different sides of the kink.

    ``oracle_run(overrides) -> (result dict, ActTape)`` runs the float64 oracle with the given mask overrides.  ``ref`` is
    what ``got`` is compared with: the oracle's own unperturbed result (default) or a golden fixture of the reference.

    1. Plain comparison.  Nothing above ``tol``: done, zero flips.
    2. Otherwise the elements whose float64 pre-activation lies within ``margin`` of the kink, relative to the RMS of
       their activation site (spectra have maxima ~100x their RMS, so the maximum is no yardstick for an absolute
       rounding error), are the only ones a rounding difference can flip.  Each candidate is flipped alone in an oracle
       replay; the change of all gradients it causes, D_e, is exact and the flips superpose (masks are piecewise constant
       and a pre-activation of ~0 leaves every forward value alone).
    3. Integer flip indicators c_e (0 / 1 against the oracle; -1 / 0 / 1 against a fixture, whose own FP32 run may have
       flipped too) are fitted to the gradient differences; the gradients corrected by the detected flips must then meet
       the tolerance everywhere.  Forward values are never corrected.  Margins are tried from tight to wide.

    ``noise_floor``: per-key errors of the REFERENCE's own FP32 arithmetic against the same float64 oracle (the FP32 CPU
    oracle run through flip_aware_errors).  A whole network at batch 1-2 through training-mode BatchNorm is ill-conditioned:
    there the reference itself sits 1e-4..5e-3 from exact on some tensors (the weight gradient of the 128x128 Fourier unit
    is a sum over 8320 bins with ~1e5 cancellation), and asking more of the product than the arithmetic can give would be
    a statement about the test, not the kernels.  With it, the bound per tensor is max(tol, 4 x the reference's own error):
    the tensor-core path keeps ~22 significant bits per product (3xTF32 with a round-to-nearest split, ffc_common.cuh)
    where FP32 keeps 24, i.e. up to 4x the rounding error of the reference on the same ill-conditioned sum.
    Returns (errs, n_flips)."""
    out_tol = tol if out_tol is None else out_tol
    errs, flips, plain = flip_aware_errors(got, oracle_run, ref, margins, max_flips, good=tol, cache=cache)
    is_grad = lambda k: k.startswith(("din", "grad/"))

    def lim(k):
        base = tol if is_grad(k) else out_tol
        return max(base, 4.0 * noise_floor.get(k, 0.0)) if noise_floor else base
    bad_fwd = {k: v for k, v in errs.items() if not is_grad(k) and not v <= lim(k)}
    assert not bad_fwd, f"{what}: forward values above tolerance: {bad_fwd}"
    bad = {k: (v, lim(k)) for k, v in errs.items() if is_grad(k) and not v <= lim(k)}
    assert not bad, (f"{what}: gradients above tolerance that no activation-mask flip explains (error, bound): {bad}; "
                     f"{flips} flips fitted; before fitting: { {k: plain[k] for k in bad} }")
    return errs, flips
