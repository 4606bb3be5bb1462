"""Image comparison for path-matched renders (GPU float pipeline vs fp64 oracle on the same Philox streams).

Both sides walk the same paths, so almost every pixel agrees to ~1e-5.  A float-vs-double rounding difference can
flip a discrete decision (a roulette draw at its threshold, a silhouette hit) for a handful of paths; such a pixel then
differs by a whole path contribution, which would dominate a plain RMSE.  The check therefore bounds the FRACTION of
disagreeing pixels and the RMSE over the agreeing ones."""
import numpy as np


def images_match(rgb, ref, pixel_tol=2e-3, max_bad_fraction=3e-3, rmse_tol=5e-4):
    rgb = np.asarray(rgb, np.float64); ref = np.asarray(ref, np.float64)
    mean = max(float(ref.mean()), 1e-12)
    diff = np.abs(rgb - ref).max(axis=-1)
    bad = diff > pixel_tol * (ref.max(axis=-1) + mean)
    frac = float(bad.mean())
    good = ~bad
    rmse = float(np.sqrt(((rgb - ref)[good] ** 2).mean()) / mean) if good.any() else 0.0
    ok = frac <= max_bad_fraction and rmse <= rmse_tol
    return ok, {"bad_pixel_fraction": frac, "rel_rmse_of_matching_pixels": rmse, "n_bad": int(bad.sum())}


def plog(**kw):
    """Append one JSON line of measured parity figures to $DSRT_PARITY_LOG (profiles/r2_parity.md is built from it)."""
    import json
    import os
    p = os.environ.get("DSRT_PARITY_LOG")
    if p:
        with open(p, "a") as f:
            f.write(json.dumps(kw) + "\n")
