"""Shards that do not fit the device: ranges of tiles under a tile budget, and position windows through the two
window slots of the C ABI (msnv_window_begin / _add_sample / _run). Both must give exactly the hits of the
resident whole-shard run."""
import os

import numpy as np
import pytest

from metasnv_b200 import harness as H

pytestmark = pytest.mark.gpu

FIELDS = ("pos", "pop_mask", "ind_mask", "cov", "allele", "total")


def _same_hits(a, b):
    for f in FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


def _concat(hs):
    class R:
        pass
    r = R()
    for f in FIELDS:
        setattr(r, f, np.concatenate([getattr(h, f) for h in hs], axis=0))
    return r


@pytest.mark.parametrize("preset,scale,samples", [("c2", 0.02, 24), ("c4", 0.004, 6), ("c3", 0.002, 40)])
def test_tile_budget_ranges_give_identical_hits(preset, scale, samples, built, monkeypatch):
    from metasnv_b200 import abi
    desc = H.describe(preset, scale, samples)
    with abi.Context(0) as ctx:
        P, first = ctx.shard_synth(desc)
        if first >= 0:
            ctx.shard_mask_position(first)
        whole = ctx.shard_run()
        t = ctx.timings()
        assert t["n_ranges"] == 1 and whole.n_hits > 0
        # a budget of a few tiles' worth of count planes: the run is cut into many ranges
        monkeypatch.setenv("MSNV_TILE_BUDGET_MB", "%.4f" % (samples * 12 * 3.5 / 1024))
        split = ctx.shard_run()
        assert ctx.timings()["n_ranges"] > 2
        _same_hits(whole, split)
        monkeypatch.delenv("MSNV_TILE_BUDGET_MB")
        _same_hits(whole, ctx.shard_run())


@pytest.mark.parametrize("preset,scale,samples,win_tiles", [("c2", 0.02, 24, 16), ("c1", 0.05, 20, 7), ("c4", 0.004, 6, 3)])
def test_position_windows_give_identical_hits(preset, scale, samples, win_tiles, built):
    """The shard cut into windows of `win_tiles` tiles; each window gets, per sample, only the reads that reach it
    (abi.window_slice); uploads of window k+1 are queued before window k runs (the two slots overlap)."""
    from metasnv_b200 import abi
    desc = H.describe(preset, scale, samples)
    with abi.Context(0) as ctx:
        P, first = ctx.shard_synth(desc)
        S = desc["n_samples"]
        exported = [ctx.export_sample(s) for s in range(S)]
        ref = ctx.export_ref(P)
        if first >= 0:
            ctx.shard_mask_position(first)
        whole = ctx.shard_run()
    T = abi.TILE
    bounds = list(range(0, P, win_tiles * T)) + [P]
    wins = list(zip(bounds[:-1], bounds[1:]))
    with abi.Context(0) as ctx:
        ctx.shard_begin(S, ref)
        if first >= 0:
            ctx.shard_mask_position(first)

        def upload(k):
            lo, hi = wins[k]
            ctx.window_begin(k & 1, lo, hi)
            for s, e in enumerate(exported):
                if e["pos"].size:
                    w = abi.window_slice(e, lo, hi)
                    if w is not None:
                        ctx.window_add_sample(k & 1, s, w)
        parts = []
        upload(0)
        for k in range(len(wins)):
            if k + 1 < len(wins):
                upload(k + 1)
            h = ctx.window_run(k & 1)
            assert np.all((h.pos >= wins[k][0]) & (h.pos < wins[k][1]))
            parts.append(h)
    assert len(wins) > 2
    _same_hits(whole, _concat(parts))
