"""Helpers for the -m gpu parity tests: feed the same Scene to the CUDA library and to the oracle."""
from __future__ import annotations

import numpy as np

import physkit_b200 as pk


def make_context(scene, max_pairs, mode=pk.MODE_QUERY, **kw):
    nh = sum(len(s[1]) for s in scene.shapes if s[0] == "hull")
    wid = kw.pop("world_id_array", None)
    ctx = pk.Context(max(scene.n, 2), max_pairs, mode=mode, max_shapes=max(len(scene.shapes), 1),
                     max_hull_vertices=nh + 8, **kw)
    ids = ctx.add_shapes(scene.shapes)
    assert ids == list(range(len(scene.shapes)))
    ctx.resize(scene.n)
    ctx.upload(scene.pos, scene.quat, scene.disp, scene.shape_id, scene.flags, wid)
    return ctx


def contacts_equal_bitwise(gpu_rec, hit_gpu, hit_ref, out_ref):
    """gpu_rec: structured pk_contact array per pair; out_ref: [n,10] from the oracle."""
    assert np.array_equal(hit_gpu, hit_ref), f"hit flags differ at {np.nonzero(hit_gpu != hit_ref)[0][:10]}"
    m = hit_ref.astype(bool)
    got = np.concatenate([gpu_rec["normal"], gpu_rec["world_a"], gpu_rec["world_b"], gpu_rec["depth"][:, None]], axis=1)
    # bit-exact: compare the raw 64-bit patterns (−0.0 vs +0.0 would differ here too)
    a = got[m].view(np.uint64)
    b = np.ascontiguousarray(out_ref[m]).view(np.uint64)
    if not np.array_equal(a, b):
        bad = np.nonzero((a != b).any(axis=1))[0]
        raise AssertionError(f"{len(bad)} of {m.sum()} contacts differ bitwise; first: pair {np.nonzero(m)[0][bad[0]]}\n"
                             f"gpu {got[m][bad[0]]}\nref {out_ref[m][bad[0]]}")


class GpuGjk:
    """gjk_epa(a, b) through pk_gjk_epa_batch on one long-lived context (shapes are cached)."""

    def __init__(self):
        self.ctx = pk.Context(2, 16, mode=pk.MODE_QUERY, max_shapes=4096, max_hull_vertices=1 << 20)
        self.cache = {}
        self.ctx.resize(2)

    def _shape(self, spec):
        key = id(spec[1]) if spec[0] == "hull" else spec
        if key not in self.cache:
            self.cache[key] = (self.ctx.add_shape(spec), spec)  # keep spec alive so id() stays unique
        return self.cache[key][0]

    def __call__(self, a, b):
        ia, ib = self._shape(a[0]), self._shape(b[0])
        self.ctx.upload([a[1], b[1]], [a[2], b[2]], None, [ia, ib], [2, 2])
        hit, out = self.ctx.gjk_epa_batch([0], [1])
        if not hit[0]:
            return None
        o = out[0]
        return dict(normal=o["normal"].copy(), world_a=o["world_a"].copy(), world_b=o["world_b"].copy(), depth=float(o["depth"]))

    def close(self):
        self.ctx.close()
