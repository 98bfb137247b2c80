"""Flip-aware parity report shared by the GPU parity tests.

The rasterizer specification has hard decisions -- alpha < 1/255, T < 1e-4,
alpha clamp at 0.99, integer radius, tile-rect truncation, near/front culls, clamp edges -- where two
correct fp32 evaluations can land on different sides.  Instead of tolerating an unexplained share of
outliers, the float64 oracle is evaluated three times: as specified and with EVERY hard threshold
shifted by +-SHIFT (relative; ~20x the fp32 rounding error of the compared quantities).  An element
whose value moves under the shift is FLIP-PRONE: its value provably hinges on a comparison within
2e-5 of its threshold.  The checks:

  (1) the 2-norm of the error over ALL elements that are not flip-prone (nothing else removed) must be
      <= tol * ||ref||.  With `floor` (the same formulas evaluated in fp32 by the C oracle) tol becomes
      max(tol, 2 x the fp32 oracle's own error): no worse than fp32 evaluation noise.
  (2) element-wise, EVERY element (flip-prone or not) must satisfy
          |got - ref| <= tol * scale + 1.5 * D + 3 * E32          (scale = max |ref|)
      D = the change the threshold shift itself produces at that element (a flip can only move it to the
      other side of the same decision), E32 = |fp32 oracle - fp64 oracle| at that element (0 without a
      floor): an outlier is accepted only where the oracle itself shows the sensitivity, and only up
      to that magnitude.
      The tail of plain fp32 evaluation noise may break (2) at no more than `max_tail_frac` (5e-5) of
      the elements, and then by no more than `tail_mult` (10) x tol: bounded in share AND magnitude.
  (3) the flip-prone share is bounded (`max_flip_frac`), so (1)-(2) cover almost everything.
"""
import torch

TOL = 1e-4
SHIFT = 2e-5


def flip_report(name, got, ref0, refp, refm, *, tol=TOL, floor=None, max_flip_frac=0.03, max_tail_frac=5e-5,
                tail_mult=10.0, quiet=False):
    got, ref0, refp, refm = [t.detach().double().cpu().reshape(-1) for t in (got, ref0, refp, refm)]
    assert got.shape == ref0.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref0.shape)}"
    if ref0.numel() == 0:
        return True
    scale = max(ref0.abs().max().item(), 1e-300)
    rnorm = max(ref0.norm().item(), 1e-300)
    D = torch.maximum((refp - ref0).abs(), (refm - ref0).abs())
    D = torch.maximum(D, (refp - refm).abs())
    prone = D > 0.25 * tol * scale
    stable = ~prone
    err = (got - ref0).abs()
    tol_eff, floor_l2, E32 = tol, None, torch.zeros_like(err)
    if floor is not None:
        E32 = (floor.detach().double().cpu().reshape(-1) - ref0).abs()
        floor_l2 = (E32 * stable).norm().item() / rnorm
        tol_eff = max(tol, 2.0 * floor_l2)
    l2 = (err * stable).norm().item() / rnorm
    emax = (err * stable).max().item() / scale
    # element-wise: tol*scale, plus what the threshold shift itself does to the element, plus (with a
    # floor) what plain fp32 evaluation of the same formulas does to it
    tail = err > tol_eff * scale + 1.5 * D + 3.0 * E32
    tail_frac = tail.double().mean().item()
    # the tail of plain fp32 evaluation noise (sharp splats amplify the ~1e-4 px error of the projected
    # centre): at most max_tail_frac of the elements, each within tail_mult x tol
    violations = int((err > tail_mult * tol_eff * scale + 1.5 * D).sum())
    frac = prone.double().mean().item()
    ok = (l2 <= tol_eff) and violations == 0 and tail_frac <= max_tail_frac and frac <= max_flip_frac
    if not quiet:
        extra = "" if floor_l2 is None else f" fp32-oracle l2={floor_l2:.2e}"
        print(f"  {name:12s} stable: l2_rel={l2:.2e} max_rel={emax:.2e} | flip-prone frac={frac:.2e} | tail frac={tail_frac:.1e} "
              f"violations={violations} (tol {tol_eff:.1e}{extra}) {'ok' if ok else 'FAIL'}")
    return ok


def int_flip_report(name, got, ref0, refp, refm, max_flip_frac=0.02):
    """integer outputs (count, radii): exact outside the flip-prone set, within the shifted range inside"""
    got, ref0, refp, refm = [t.detach().long().cpu().reshape(-1) for t in (got, ref0, refp, refm)]
    if ref0.numel() == 0:
        return True
    lo = torch.minimum(torch.minimum(refp, refm), ref0)
    hi = torch.maximum(torch.maximum(refp, refm), ref0)
    prone = hi > lo
    bad_stable = int(((got != ref0) & ~prone).sum())
    span = hi - lo
    bad_prone = int((((got < lo - span) | (got > hi + span)) & prone).sum())
    frac = prone.double().mean().item()
    # a stable mismatch is a comparison decided by fp32 noise beyond the shift: <= 5e-5 of the elements, off by one
    off_by_more = int((((got - ref0).abs() > 1) & ~prone).sum())
    ok = bad_stable <= 5e-5 * ref0.numel() and off_by_more == 0 and bad_prone == 0 and frac <= max_flip_frac
    print(f"  {name:12s} stable: mismatches={bad_stable} | flip-prone: frac={frac:.2e} out-of-range={bad_prone} "
          f"{'ok' if ok else 'FAIL'}")
    return ok
