"""Spatial pruning of the tensor-core Chamfer filter (csrc/chamfer_prep.cu + chamfer_tc.cu): the targets are
Morton-sorted and whole 128 x 256 stages are skipped when box gaps exceed known-achievable distances.  Results must not
change by a bit - min, arg-min, and the first-ORIGINAL-index tie rule (torch.min, chamfer_distance.py:22-23) although
the targets are permuted internally - and on primitive-shaped inputs a large share of the stages must really be skipped."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vpn():
    assert torch.cuda.is_available()
    import vpn_b200
    return vpn_b200


def primitive_scene(gen, b, k, n, m, lattice=False):
    """K small cuboid-surface patches of N consecutive rows each (the layout of train.py:119) against M targets on the
    surface of a few boxes, in RANDOM order (ShapeNet samples are unordered)."""
    centres = (torch.rand(b, k, 1, 3, generator=gen) - 0.5) * 0.8
    ext = torch.rand(b, k, 1, 3, generator=gen) * 0.1 + 0.02
    u = torch.rand(b, k, n, 3, generator=gen) * 2 - 1
    axis = torch.arange(n)[None, None, :] * 3 // n                      # face-major within a primitive, like the sampler
    u.scatter_(3, axis[..., None].expand(b, k, n, 1), 1.0)
    p1 = (centres + u * ext).reshape(b, k * n, 3)
    tc = (torch.rand(b, 6, 3, generator=gen) - 0.5) * 0.7
    th = torch.rand(b, 6, 3, generator=gen) * 0.12 + 0.03
    w = torch.randint(0, 6, (b, m), generator=gen)
    t = torch.rand(b, m, 3, generator=gen) * 2 - 1
    ax = torch.randint(0, 3, (b, m), generator=gen)
    t.scatter_(2, ax[..., None], (torch.randint(0, 2, (b, m, 1), generator=gen).float() * 2 - 1))
    bi = torch.arange(b)[:, None]
    p2 = tc[bi, w] + t * th[bi, w]
    if lattice:                                                         # exact ties between distinct targets and duplicates
        p1 = torch.round(p1 * 32) / 32
        p2 = torch.round(p2 * 32) / 32
    return p1.contiguous(), p2.contiguous()


def stats(vpn, p1, p2, impl=5):
    from vpn_b200 import _lib
    lib = _lib.load()
    b, p, _ = p1.shape
    m = p2.shape[1]
    nb = ctypes.c_size_t(0)
    _lib.check(lib.vpn_chamfer_workspace_bytes(b, p, m, impl, ctypes.byref(nb)), "ws")
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    out = [torch.empty(b, p, device="cuda"), torch.empty(b, p, dtype=torch.int32, device="cuda"),
           torch.empty(b, m, device="cuda"), torch.empty(b, m, dtype=torch.int32, device="cuda")]
    _lib.check(lib.vpn_chamfer_fwd(_lib.ptr(p1), _lib.ptr(p2), *[_lib.ptr(o) for o in out], b, p, m, _lib.ptr(ws), nb.value, impl,
                                   _lib.stream_ptr(p1.device)), "fwd")
    st, sk = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    _lib.check(lib.vpn_chamfer_prune_stats(_lib.ptr(ws), b, p, m, impl, ctypes.byref(st), ctypes.byref(sk), _lib.stream_ptr(p1.device)), "stats")
    return out, int(st.value), int(sk.value)


@pytest.mark.parametrize("b,k,n,m,lattice", [
    (2, 16, 1024, 4096, False),         # C2-like, reduced
    (2, 16, 1024, 4096, True),          # ties between distinct targets: the lowest ORIGINAL index must win
    (1, 7, 700, 3001, False),           # nothing a multiple of 128
    (1, 32, 512, 16384, False),         # largest sample the in-shared-memory sort takes
    (1, 8, 1024, 16385, False),         # one more: unsorted sweep (identity permutation), still exact
    (3, 4, 128, 129, True),
])
def test_pruned_filter_is_bit_exact(vpn, c_oracle, b, k, n, m, lattice):
    gen = torch.Generator().manual_seed(1000 + m)
    p1, p2 = primitive_scene(gen, b, k, n, m, lattice)
    want = c_oracle(p1.numpy(), p2.numpy())
    (m1, i1, m2, i2), stages, skipped = stats(vpn, p1.cuda(), p2.cuda())
    np.testing.assert_array_equal(i1.cpu().numpy().astype(np.int64), want[1], err_msg="idx1")
    np.testing.assert_array_equal(i2.cpu().numpy().astype(np.int64), want[3], err_msg="idx2")
    np.testing.assert_array_equal(m1.cpu().numpy(), want[0]); np.testing.assert_array_equal(m2.cpu().numpy(), want[2])
    assert stages > 0 and 0 <= skipped <= stages


def test_pruning_skips_most_of_a_primitive_scene_and_can_be_switched_off(vpn):
    """On the C2 layout (16 cuboid patches x 4096 rows vs 8192 shuffled surface targets) more than 40 % of the stages go;
    with vpn_set_tuning("tc_prune", 2) none do and every output is identical."""
    from vpn_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(5)
    p1, p2 = primitive_scene(gen, 2, 16, 4096, 8192)
    p1, p2 = p1.cuda(), p2.cuda()
    on, stages, skipped = stats(vpn, p1, p2)
    assert skipped > 0.4 * stages, (skipped, stages)
    assert lib.vpn_set_tuning(b"tc_prune", 2) == 0
    try:
        off, stages_off, skipped_off = stats(vpn, p1, p2)
    finally:
        lib.vpn_set_tuning(b"tc_prune", 0)
    assert skipped_off == 0 and stages_off == stages
    for a, c in zip(on, off):
        assert torch.equal(a, c)


def test_shuffled_rows_are_resorted_and_stay_exact(vpn, c_oracle):
    """Rows in random order have no compact 128-row blocks as they arrive; the prep pass Morton-sorts every 4096-row
    segment (chamfer_prep.cu), so stages still prune, and the results - returned in the ORIGINAL row order, ties to the
    first original index - stay exact.  A cloud shorter than one segment and one with a ragged last segment included."""
    gen = torch.Generator().manual_seed(9)
    for (k, n, m) in ((8, 1024, 2048), (3, 1000, 1536), (9, 1000, 2048)):
        p1, p2 = primitive_scene(gen, 1, k, n, m, lattice=(k == 9))
        p1 = p1[:, torch.randperm(p1.shape[1], generator=gen)].contiguous()
        want = c_oracle(p1.numpy(), p2.numpy())
        (m1, i1, m2, i2), stages, skipped = stats(vpn, p1.cuda(), p2.cuda())
        np.testing.assert_array_equal(i1.cpu().numpy().astype(np.int64), want[1]); np.testing.assert_array_equal(i2.cpu().numpy().astype(np.int64), want[3])
        np.testing.assert_array_equal(m1.cpu().numpy().view(np.int32), want[0].view(np.int32))
        np.testing.assert_array_equal(m2.cpu().numpy().view(np.int32), want[2].view(np.int32))
        assert 0 < skipped < stages


def test_compact_blocks_keep_their_order(vpn, c_oracle):
    """Mesh vertices arrive one small primitive per 128-row block (train_gcn.py:127-130): the segment sort must not make
    the blocks less compact - the prep pass keeps the natural order there, and the pruning stays as good as the blocks."""
    gen = torch.Generator().manual_seed(10)
    b, k, nv, m = 1, 64, 128, 4096
    centres = (torch.rand(b, k, 1, 3, generator=gen) - 0.5) * 0.9
    d = torch.randn(b, k, nv, 3, generator=gen)
    p1 = (centres + 0.02 * d / d.norm(dim=-1, keepdim=True)).reshape(b, k * nv, 3).contiguous()
    p2 = (torch.rand(b, m, 3, generator=gen) - 0.5).contiguous()
    want = c_oracle(p1.numpy(), p2.numpy())
    (m1, i1, m2, i2), stages, skipped = stats(vpn, p1.cuda(), p2.cuda())
    np.testing.assert_array_equal(i1.cpu().numpy().astype(np.int64), want[1]); np.testing.assert_array_equal(i2.cpu().numpy().astype(np.int64), want[3])
    assert skipped > 0.6 * stages          # natural order: ~0.65-0.7 (ideal 0.72); the segment sort would leave < 0.57


def test_probe_knobs_keep_results_exact(vpn, c_oracle):
    """The probe knobs of the filter and the prep pass change how the work is done, never a result: epilogue reduction
    on the FP16 pipe for 0 / 2 / 4 units of a block (tc_hunits), reproducible sort permutation (prep_deterministic),
    narrower / wider pruning bounds.  With prep_deterministic the pruning statistics repeat exactly."""
    from vpn_b200 import _lib
    lib = _lib.load()
    gen = torch.Generator().manual_seed(21)
    p1, p2 = primitive_scene(gen, 2, 6, 1500, 3000, lattice=True)
    want = c_oracle(p1.numpy(), p2.numpy())
    try:
        for knobs in ({b"tc_hunits": 1}, {b"tc_hunits": 3}, {b"tc_hunits": 5}, {b"prep_deterministic": 1},
                      {b"prep_near_rows": 2, b"prep_reps_rows": 1, b"prep_near_cols": 4, b"prep_reps_cols": 2}, {b"prep_probe": 1}):
            for k, v in knobs.items():
                _lib.check(lib.vpn_set_tuning(k, v), "tuning")
            (m1, i1, m2, i2), stages, skipped = stats(vpn, p1.cuda(), p2.cuda())
            np.testing.assert_array_equal(i1.cpu().numpy().astype(np.int64), want[1]); np.testing.assert_array_equal(i2.cpu().numpy().astype(np.int64), want[3])
            np.testing.assert_array_equal(m1.cpu().numpy().view(np.int32), want[0].view(np.int32))
            np.testing.assert_array_equal(m2.cpu().numpy().view(np.int32), want[2].view(np.int32))
            if b"prep_deterministic" in knobs:                  # continuous coordinates: no cell above the in-order cap
                q1, q2 = primitive_scene(gen, 2, 6, 1500, 3000)
                first = stats(vpn, q1.cuda(), q2.cuda())
                again = stats(vpn, q1.cuda(), q2.cuda())
                assert (again[1], again[2]) == (first[1], first[2])
                for a, c in zip(first[0], again[0]):
                    assert torch.equal(a, c)
            for k in knobs:
                lib.vpn_set_tuning(k, 0)
    finally:
        for k in (b"tc_hunits", b"prep_deterministic", b"prep_near_rows", b"prep_reps_rows", b"prep_near_cols", b"prep_reps_cols", b"prep_probe"):
            lib.vpn_set_tuning(k, 0)
