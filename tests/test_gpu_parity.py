"""GPU parity tests: every call goes product API -> ctypes -> C ABI -> sm_100a kernel, and is compared
with the CPU oracle (oracle/vpn_oracle.py, oracle/chamfer_oracle.c) and with the golden vectors the
reference itself produced (tests/golden/).  Bars: arg-min indices and integer outputs bit-exact;
floating point within 1e-4 relative (BASELINE.json north_star), tolerance written at each assert.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star: "losses, silhouettes and gradients within 1e-4 relative in FP32"


@pytest.fixture(scope="module")
def vpn():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import vpn_b200
    return vpn_b200


@pytest.fixture(scope="module")
def O():
    from oracle import vpn_oracle
    return vpn_oracle


def C(a):
    return torch.as_tensor(np.asarray(a)).cuda()


def close(a, b, rtol=RTOL, atol=1e-6, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


def same(a, b, what=""):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_array_equal(a, b, err_msg=what)


# ------------------------------------------------------------------------------------------------
# transform
# ------------------------------------------------------------------------------------------------
def test_transform_golden(vpn, golden):
    g = golden
    import modules.transform as mt
    pts, q, t = C(g["in_tf_points"]), C(g["in_tf_q"]), C(g["in_tf_t"])
    close(mt.rotate_points(pts, q), g["ref_tf_rotate"], atol=1e-6)
    close(mt.transform_points(pts, q, t), g["ref_tf_transform"], atol=1e-6)
    d, e, a, ang = (C(g["in_tf_" + k]) for k in ("dists", "elevs", "azims", "angles"))
    close(mt.view_to_obj_points(pts, d, e, a, ang), g["ref_tf_view_to_obj"], atol=2e-6)
    close(mt.obj_to_view_points(pts, d, e, a), g["ref_tf_obj_to_view"], atol=2e-6)
    close(mt.rotate_points_forward_x_axis(pts, ang), g["ref_tf_rotate_x"], atol=1e-6)
    close(mt.translate_points(pts, t), g["in_tf_points"] + g["in_tf_t"][:, None, :], atol=0, rtol=0)
    # gradients
    pts, q, t = (x.clone().requires_grad_() for x in (pts, q, t))
    (mt.transform_points(pts, q, t) * C(g["in_tf_upstream"])).sum().backward()
    close(pts.grad, g["ref_tf_grad_points"], atol=1e-6)
    close(q.grad, g["ref_tf_grad_q"], atol=2e-5)
    close(t.grad, g["ref_tf_grad_t"], atol=1e-6)
    pg = C(g["in_tf_points"]).requires_grad_()
    (mt.view_to_obj_points(pg, d, e, a, ang) * C(g["in_tf_upstream"])).sum().backward()
    close(pg.grad, g["ref_tf_view_to_obj_grad_points"], atol=2e-6)


def test_transform_random_vs_oracle(vpn, O):
    gen = torch.Generator().manual_seed(7)
    for b, n in ((1, 1), (3, 5), (2, 1000), (5, 4099)):
        pts = torch.randn(b, n, 3, generator=gen); q = torch.randn(b, 4, generator=gen) * 2; t = torch.randn(b, 3, generator=gen)
        w = torch.randn(b, n, 3, generator=gen)
        po, qo, to = (x.clone().requires_grad_() for x in (pts, q, t))
        (O.transform_points(po, qo, to) * w).sum().backward()
        pc, qc, tc = (x.cuda().requires_grad_() for x in (pts, q, t))
        out = vpn.transform_points(pc, qc, tc)
        close(out, O.transform_points(pts, q, t), atol=2e-6)
        (out * w.cuda()).sum().backward()
        close(pc.grad, po.grad, atol=2e-6)
        close(tc.grad, to.grad, atol=1e-5 * max(1, n ** 0.5))
        close(qc.grad, qo.grad, rtol=2e-4, atol=2e-5 * max(1.0, float(qo.grad.abs().max())))


# ------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------
def test_sphere_sampling_golden(vpn, golden):
    g = golden
    v, q, t = (C(g["in_sp_" + k]).requires_grad_() for k in ("v", "q", "t"))
    u = torch.cat([C(g["in_sp_ue"]), C(g["in_sp_ua"])], dim=2)[:, None]       # (B,1,N,2)
    out = vpn.sample_primitives("sphere", v[:, None], q[:, None], t[:, None], u)
    close(out, g["ref_sp_points"], atol=1e-6)
    (out * C(g["in_sp_upstream"])).sum().backward()
    close(v.grad, g["ref_sp_grad_v"], atol=1e-5); close(q.grad, g["ref_sp_grad_q"], atol=1e-5); close(t.grad, g["ref_sp_grad_t"], atol=1e-5)


def test_cuboid_sampling_golden(vpn, golden):
    g = golden
    v, q, t = (C(g["in_cb_" + k]).requires_grad_() for k in ("v", "q", "t"))
    u = C(g["in_cb_u"])[:, None]
    same(vpn.cuboid_face_counts(v.detach(), u.shape[2]), g["ref_cb_counts"], "face counts must be bit-exact")
    same(vpn.cuboid_face_counts(C(g["in_cb_counts_v"]), 1000), g["ref_cb_counts_1000"])
    same(vpn.cuboid_face_counts(C(g["in_cb_counts_v"]), 4096), g["ref_cb_counts_4096"])
    out = vpn.sample_primitives("cuboid", v[:, None], q[:, None], t[:, None], u)
    close(out, g["ref_cb_points"], atol=1e-6)
    (out * C(g["in_cb_upstream"])).sum().backward()
    close(v.grad, g["ref_cb_grad_v"], atol=1e-5); close(q.grad, g["ref_cb_grad_q"], atol=1e-5); close(t.grad, g["ref_cb_grad_t"], atol=1e-5)


@pytest.mark.parametrize("kind", ["sphere", "cuboid"])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 7), (2, 16, 128), (3, 2, 1026), (1, 2, 4096)])
def test_sampling_random_vs_oracle(vpn, O, kind, shape):
    b, k, n = shape
    v, q, t = O.synthetic_primitives(b, k, seed=11)
    gen = torch.Generator().manual_seed(5)
    u = torch.rand(b, k, n, 2 if kind == "sphere" else 3, generator=gen)
    w = torch.randn(b, k * n, 3, generator=gen)
    vo, qo, to = (x.clone().requires_grad_() for x in (v, q, t))
    ref = O.sample_predict_points(kind, vo, qo, to, u)
    (ref * w).sum().backward()
    vc, qc, tc = (x.cuda().requires_grad_() for x in (v, q, t))
    out = vpn.sample_primitives(kind, vc, qc, tc, u.cuda())
    assert out.shape == (b, k * n, 3)
    close(out, ref, atol=1e-6)
    (out * w.cuda()).sum().backward()
    scale = max(1.0, n ** 0.5)
    close(vc.grad, vo.grad, rtol=2e-4, atol=2e-5 * scale)
    close(tc.grad, to.grad, rtol=2e-4, atol=2e-5 * scale)
    close(qc.grad, qo.grad, rtol=2e-4, atol=2e-5 * scale)


def test_mesh_vertices_golden(vpn, golden, golden_templates):
    g = golden
    from vpn_b200 import templates
    v, q, t = C(g["in_ms_v"]), C(g["in_ms_q"]), C(g["in_ms_t"])
    for name in ("sphere", "cuboid"):
        tv, tf = templates.template(name, "cuda")
        out = vpn.mesh_vertices(tv, v[:, None], q[:, None], t[:, None])
        close(out, g[f"ref_ms_{name}_vertices"], atol=1e-6)
        same(tf, golden_templates[name + "_faces"])
    import modules.meshing as mm
    sm, cm = mm.Meshing.sphere_meshing(v, q, t), mm.Meshing.cuboid_meshing(v, q, t)
    comp = mm.Meshing.compose_meshes([sm[0], cm[0], sm[1]])
    close(comp.vertices, g["ref_ms_compose_vertices"], atol=1e-6)
    same(comp.faces, g["ref_ms_compose_faces"])
    assert comp.faces.dtype == torch.int64


# ------------------------------------------------------------------------------------------------
# Chamfer
# ------------------------------------------------------------------------------------------------
IMPLS = {"generic": 1, "tiled_exact": 2, "tiled_fma": 3, "tiled_expand": 4, "tiled_tc": 5, "auto": 0}


def run_nn(vpn, p1, p2, impl):
    m1, i1, m2, i2 = vpn.chamfer_nn(torch.as_tensor(p1).cuda(), torch.as_tensor(p2).cuda(), impl)
    return m1.cpu().numpy(), i1.cpu().numpy().astype(np.int64), m2.cpu().numpy(), i2.cpu().numpy().astype(np.int64)


def test_chamfer_golden(vpn, golden):
    g = golden
    m1, i1, m2, i2 = run_nn(vpn, g["in_cd_p1"], g["in_cd_p2"], 1)
    same(i1, g["ref_cd_idx1"], "idx1 bit-exact vs reference torch.min"); same(i2, g["ref_cd_idx2"])
    # golden values carry torch-CPU's VML sqrt (<= 1 ulp off); IEEE values are checked against the oracle below
    close(m1, g["ref_cd_min1"], rtol=1.3e-7, atol=0); close(m2, g["ref_cd_min2"], rtol=1.3e-7, atol=0)
    import modules.loss as ml
    p1, p2 = C(g["in_cd_p1"]).requires_grad_(), C(g["in_cd_p2"]).requires_grad_()
    loss = ml.ChamferDistanceLoss()(p1, p2)
    close(loss, g["ref_cd_loss"], atol=0)
    loss.backward()
    close(p1.grad, g["ref_cd_grad_p1"], atol=1e-8); close(p2.grad, g["ref_cd_grad_p2"], atol=1e-8)
    close(ml.ChamferDistanceLoss()(p1.detach(), p2.detach(), each_batch=True), g["ref_cd_loss_each"], atol=0)
    close(ml.ChamferDistanceLoss()(p1.detach(), p2.detach(), w1=0.5, w2=1.0), g["ref_cd_loss_w"], atol=0)
    tr = C(g["in_vd_translates"])
    close(ml.VPDiverseLoss()([tr[:, i] for i in range(tr.shape[1])], C(g["in_vd_gt"])), g["ref_vd_loss"], atol=0)


def adversarial_clouds(name, b, p, m, gen):
    r = lambda *s: torch.rand(*s, generator=gen)
    if name == "uniform":
        return r(b, p, 3) - 0.5, r(b, m, 3) - 0.5
    if name == "lattice":          # many exact ties in d and in sqrt(d)
        return torch.floor(r(b, p, 3) * 6) / 4, torch.floor(r(b, m, 3) * 6) / 4 + 0.125
    if name == "duplicates":       # every target appears ~4 times -> first index must win
        base = r(b, (m + 3) // 4, 3)
        return r(b, p, 3), base.repeat(1, 4, 1)[:, :m].contiguous()
    if name == "offset":           # far from the origin: differences lose low bits, near ties abound
        return r(b, p, 3) * 0.05 + 100.0, r(b, m, 3) * 0.05 + 100.0
    if name == "identical":        # all distances zero
        return torch.zeros(b, p, 3) + 0.25, torch.zeros(b, m, 3) + 0.25
    if name == "surface":          # clustered like primitive samples vs shape samples
        c = r(b, 1, 3)
        return c + 0.05 * torch.nn.functional.normalize(torch.randn(b, p, 3, generator=gen), dim=2), \
            c + 0.05 * torch.nn.functional.normalize(torch.randn(b, m, 3, generator=gen), dim=2)
    raise KeyError(name)


@pytest.mark.parametrize("impl", ["generic", "tiled_exact", "tiled_fma", "tiled_expand", "tiled_tc"])
@pytest.mark.parametrize("case", ["uniform", "lattice", "duplicates", "offset", "identical", "surface"])
def test_chamfer_nn_bit_exact(vpn, c_oracle, impl, case):
    gen = torch.Generator().manual_seed(sum(map(ord, case)))
    shapes = [(2, 1500, 700), (1, 4100, 130), (3, 1024, 128), (1, 9000, 2500)]
    if impl == "generic":
        shapes += [(2, 16, 300), (2, 1, 1), (1, 5, 3), (2, 300, 16)]
    for (b, p, m) in shapes:
        p1, p2 = adversarial_clouds(case, b, p, m, gen)
        ref = c_oracle(p1.numpy(), p2.numpy())
        got = run_nn(vpn, p1, p2, IMPLS[impl])
        for name, r_, g_ in zip(("min1", "idx1", "min2", "idx2"), ref, got):
            same(g_, r_, f"{impl}/{case}/{(b, p, m)}/{name}")


def test_chamfer_full_size_slice(vpn, c_oracle):
    """BASELINE config 2 cloud sizes (P=65536, M=8192) on a 2-sample slice, bit-exact vs the C oracle."""
    gen = torch.Generator().manual_seed(1234)
    p1, p2 = torch.rand(2, 65536, 3, generator=gen) - 0.5, torch.rand(2, 8192, 3, generator=gen) - 0.5
    ref = c_oracle(p1.numpy(), p2.numpy())
    for impl in ("auto", "tiled_tc", "tiled_exact", "tiled_fma", "tiled_expand", "generic"):
        got = run_nn(vpn, p1, p2, IMPLS[impl])
        for name, r_, g_ in zip(("min1", "idx1", "min2", "idx2"), ref, got):
            same(g_, r_, f"{impl}/{name}")


def test_chamfer_properties_full_batch(vpn):
    """Size-independent properties at BASELINE config 2's full size (B=32, P=65536, M=8192)."""
    gen = torch.Generator().manual_seed(3)
    p1, p2 = (torch.rand(32, 65536, 3, generator=gen) - 0.5).cuda(), (torch.rand(32, 8192, 3, generator=gen) - 0.5).cuda()
    m1, i1, m2, i2 = vpn.chamfer_nn(p1, p2)
    # (1) the reported minimum is the distance to the reported arg-min, recomputed with the reference's arithmetic
    bi = torch.arange(32, device="cuda")[:, None]
    d = p1 - p2[bi, i1.long()]
    d = torch.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    assert torch.equal(d, m1)
    d = p1[bi, i2.long()] - p2
    d = torch.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    assert torch.equal(d, m2)
    # (2) swapping the clouds swaps the outputs
    s1, j1, s2, j2 = vpn.chamfer_nn(p2, p1)
    assert torch.equal(s1, m2) and torch.equal(j1, i2) and torch.equal(s2, m1) and torch.equal(j2, i1)
    # (3) permuting the targets leaves min1 unchanged and idx1 still points at a target at that distance
    perm = torch.randperm(8192, generator=gen).cuda()
    p2p = p2[:, perm].contiguous()
    q1, k1, q2, k2 = vpn.chamfer_nn(p1, p2p)
    assert torch.equal(q1, m1) and torch.equal(q2, m2[:, perm])
    d = p1 - p2p[bi, k1.long()]
    d = torch.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    assert torch.equal(d, m1)
    assert (perm[k1.long()] == i1.long()).float().mean() > 0.9999
    # (4) both kernels agree
    g1, gi1, g2, gi2 = vpn.chamfer_nn(p1[:4].contiguous(), p2[:4].contiguous(), 1)
    assert torch.equal(g1, m1[:4]) and torch.equal(gi1, i1[:4]) and torch.equal(g2, m2[:4]) and torch.equal(gi2, i2[:4])


def test_chamfer_backward_vs_oracle(vpn, O):
    gen = torch.Generator().manual_seed(9)
    for (b, p, m) in ((2, 300, 200), (1, 2048, 1024)):
        p1, p2 = torch.rand(b, p, 3, generator=gen), torch.rand(b, m, 3, generator=gen)
        g1, g2 = torch.rand(b, p, generator=gen), torch.rand(b, m, generator=gen)
        m1, i1, m2, i2 = O.chamfer_nn(p1, p2)
        gp1, gp2 = O.chamfer_grad_from_nn(p1, p2, m1, i1, m2, i2, g1, g2)
        a, c = p1.cuda().requires_grad_(), p2.cuda().requires_grad_()
        o1, _, o2, _ = vpn.chamfer_nn(a, c)
        ((o1 * g1.cuda()).sum() + (o2 * g2.cuda()).sum()).backward()
        close(a.grad, gp1, atol=1e-6); close(c.grad, gp2, atol=1e-6)


@pytest.mark.parametrize("case", ["uniform", "lattice", "duplicates", "identical"])
def test_chamfer_small_first_cloud_bit_exact(vpn, c_oracle, case):
    """VP-diverse shapes (vp_diverse.py:12-18: K centres vs M targets) take the one-launch small-P kernel."""
    gen = torch.Generator().manual_seed(41 + len(case))
    for (b, p, m) in ((2, 16, 8192), (3, 1, 1), (1, 64, 257), (2, 5, 3), (1, 33, 1000), (32, 32, 2048), (2, 128, 4096), (1, 256, 300)):
        p1, p2 = adversarial_clouds(case, b, p, m, gen)
        ref = c_oracle(p1.numpy(), p2.numpy())
        got = run_nn(vpn, p1, p2, IMPLS["auto"])
        for name, r_, g_ in zip(("min1", "idx1", "min2", "idx2"), ref, got):
            same(g_, r_, f"smallp/{case}/{(b, p, m)}/{name}")


@pytest.mark.parametrize("shape", [(2, 300, 200), (1, 2048, 1024), (3, 16, 500), (2, 4096, 512)])
def test_chamfer_loss_fused_vs_oracle(vpn, O, shape):
    """ChamferDistanceLoss.forward (chamfer_distance.py:10-30) through the fused loss head/tail kernels:
    per-sample values, weights, and gradients to both clouds against autograd through the dense oracle."""
    b, p, m = shape
    gen = torch.Generator().manual_seed(p + m)
    p1, p2 = torch.rand(b, p, 3, generator=gen), torch.rand(b, m, 3, generator=gen)
    wb = torch.rand(b, generator=gen) + 0.5
    for (w1, w2) in ((1.0, 1.0), (0.5, 1.0)):
        ao, co = p1.clone().requires_grad_(), p2.clone().requires_grad_()
        ref = O.chamfer_dense(ao, co, each_batch=True, w1=w1, w2=w2)
        (ref * wb).sum().backward()
        a, c = p1.cuda().requires_grad_(), p2.cuda().requires_grad_()
        got = vpn.chamfer_distance(a, c, each_batch=True, w1=w1, w2=w2)
        close(got, ref, atol=0)
        (got * wb.cuda()).sum().backward()
        close(a.grad, ao.grad, atol=1e-4 * float(ao.grad.abs().max()))
        close(c.grad, co.grad, atol=1e-4 * float(co.grad.abs().max()))
        close(vpn.chamfer_distance(a.detach(), c.detach(), w1=w1, w2=w2), ref.mean(), atol=0)


@pytest.mark.parametrize("impl", ["tiled_tc", "tiled_expand"])
def test_chamfer_out_of_range_sample_falls_back(vpn, c_oracle, impl):
    """Coordinates too large for the centred-expansion filters flag the sample; it is redone exactly and the
    other samples of the batch are unaffected."""
    gen = torch.Generator().manual_seed(77)
    p1, p2 = torch.rand(3, 2048, 3, generator=gen), torch.rand(3, 640, 3, generator=gen)
    p1[1, 100] = torch.tensor([3e18, -2e18, 1e18])
    p2[2, 7] = torch.tensor([-3e18, 0.0, 1e17])
    ref = c_oracle(p1.numpy(), p2.numpy())
    got = run_nn(vpn, p1, p2, IMPLS[impl])
    for name, r_, g_ in zip(("min1", "idx1", "min2", "idx2"), ref, got):
        same(g_, r_, f"{impl}/fallback/{name}")


@pytest.mark.parametrize("shape", [(1, 512, 128), (2, 640, 129), (1, 16384, 2048), (2, 5000, 8192), (1, 2048 * 3 + 1, 1000)])
def test_chamfer_tc_shapes(vpn, c_oracle, shape):
    """Ragged tiles, column splits and every tile height of the tensor-core filter."""
    b, p, m = shape
    gen = torch.Generator().manual_seed(p * 3 + m)
    p1, p2 = adversarial_clouds("surface", b, p, m, gen)
    ref = c_oracle(p1.numpy(), p2.numpy())
    got = run_nn(vpn, p1, p2, IMPLS["tiled_tc"])
    for name, r_, g_ in zip(("min1", "idx1", "min2", "idx2"), ref, got):
        same(g_, r_, f"tc/{shape}/{name}")


def test_chamfer_zero_distance_gives_nan_like_reference(vpn):
    p = torch.rand(1, 8, 3).cuda().requires_grad_()
    vpn.chamfer_distance(p, p.detach().clone()).backward()
    assert torch.isnan(p.grad).all()        # sqrt'(0) * 0 = inf * 0, as autograd through the reference


# ------------------------------------------------------------------------------------------------
# end to end (train.py:105-120 + Chamfer) against the reference's own autograd
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["sphere", "cuboid"])
def test_end_to_end_golden(vpn, golden, kind):
    g = golden
    v, q, t = (C(g[f"in_e2e_{kind}_{k}"]).requires_grad_() for k in ("v", "q", "t"))
    step = vpn.PrimitiveLoss(vpn.PrimitiveLossConfig(kind=kind, l_vp_div=0.0))
    out = step(v, q, t, C(g[f"in_e2e_{kind}_u"]), C(g[f"in_e2e_{kind}_target"]))
    close(out["points"], g[f"ref_e2e_{kind}_points"], atol=1e-6)
    close(out["total"], g[f"ref_e2e_{kind}_loss"], atol=0)
    out["total"].backward()
    for k, x in (("v", v), ("q", q), ("t", t)):
        ref = g[f"ref_e2e_{kind}_grad_{k}"]
        close(x.grad, ref, atol=1e-4 * float(np.abs(ref).max()), what=f"grad_{k}")


# ------------------------------------------------------------------------------------------------
# soft silhouette (oracle = the DIB-R restatement; parity unpinned, see oracle/vpn_oracle.py)
# ------------------------------------------------------------------------------------------------
def _sil_case(O, kind, b, k, seed):
    from vpn_b200 import templates
    v, q, t = O.synthetic_primitives(b, k, seed=seed)
    t = t * 0.25
    v = v * 1.5
    tv, tf = templates.template(kind, "cpu")
    return v, q, t, tv, tf


@pytest.mark.parametrize("kind,b,k,res", [("sphere", 2, 3, 32), ("cuboid", 1, 2, 48), ("sphere", 1, 16, 64), ("sphere", 1, 2, 256)])
def test_silhouette_vs_oracle(vpn, O, kind, b, k, res):
    v, q, t, tv, tf = _sil_case(O, kind, b, k, seed=21)
    _, faces_ref = O.compose_primitive_meshes(tv, tf.long(), v, q, t)
    # the rasteriser works on coordinates x 1000 and cancels: both sides get the SAME vertices (the device's), so that
    # only the rasteriser's own arithmetic is compared (vertex parity has its own tests)
    verts = vpn.mesh_vertices(tv.cuda(), v.cuda(), q.cuda(), t.cuda()).detach().requires_grad_()
    verts_ref = verts.detach().cpu().clone().requires_grad_()
    cams = [O.look_at_camera(0.0, 0.0, 1.0) for _ in range(b)]
    rot, pos = torch.stack([c[0] for c in cams]), torch.stack([c[1] for c in cams])
    ref = O.soft_silhouette(verts_ref, faces_ref, rot, pos, res, res)
    gen = torch.Generator().manual_seed(4)
    w = torch.rand(b, res, res, generator=gen)
    (ref * w).sum().backward()
    faces = torch.cat([tf + i * tv.shape[0] for i in range(k)]).cuda()
    zero, one = torch.zeros(b).cuda(), torch.ones(b).cuda()
    r_c, p_c = vpn.look_at_cameras(zero, zero, one)
    close(r_c, rot, atol=1e-7); close(p_c, pos, atol=1e-7)
    alpha, covered, _ = vpn.soft_silhouette(verts, faces, r_c, p_c, res, res)
    # 1e-4 relative plus an absolute floor: alpha = 1 - prod(1 - p) cancels for faint pixels
    close(alpha, ref, rtol=RTOL, atol=2e-6)
    assert ((alpha.detach().cpu() == 1.0) == (ref.detach() == 1.0)).all()
    (alpha * w.cuda()).sum().backward()
    gref = verts_ref.grad
    close(verts.grad, gref, rtol=1e-3, atol=1e-4 * float(gref.abs().max()))


def test_silhouette_loss_dropin(vpn, O):
    import modules.loss as ml
    import modules.meshing as mm
    b, k, res = 2, 4, 128
    v, q, t, tv, tf = _sil_case(O, "sphere", b, k, seed=8)
    meshes_k = [mm.Meshing.sphere_meshing(v[:, i].cuda(), q[:, i].cuda(), t[:, i].cuda()) for i in range(k)]
    meshes = [mm.Meshing.compose_meshes([meshes_k[i][s] for i in range(k)]) for s in range(b)]
    gt = (torch.rand(b, 1, res, res, generator=torch.Generator().manual_seed(1)) > 0.5).float()
    dists, elevs, azims = torch.ones(b), torch.zeros(b), torch.zeros(b)
    loss = ml.SilhouetteLoss()(meshes, gt.cuda(), dists.cuda(), elevs.cuda(), azims.cuda())
    verts_ref, faces_ref = O.compose_primitive_meshes(tv, tf.long(), v, q, t)
    ref = O.silhouette_loss(verts_ref, faces_ref, gt, dists, elevs, azims)
    close(loss, ref, atol=1e-7)
    import modules.render as mr
    rgb, alpha, normals = mr.VertexRenderer.render(meshes[0], dists[0], elevs[0], azims[0])
    assert rgb.shape == (1, 128, 128, 3) and alpha.shape == (1, 128, 128, 1) and normals.shape == (1, k * 252, 3)


# ------------------------------------------------------------------------------------------------
# drop-in surface
# ------------------------------------------------------------------------------------------------
def test_sampling_dropin_consumes_reference_rng_stream(vpn, O):
    import modules.sampling as ms
    v, q, t = (x[:, 0].cuda() for x in O.synthetic_primitives(3, 1, seed=2))
    for kind in ("sphere", "cuboid"):
        torch.manual_seed(1234)
        fn = ms.Sampling.sphere_sampling if kind == "sphere" else ms.Sampling.cuboid_sampling
        out = fn(v, q, t, 200)
        torch.manual_seed(1234)
        if kind == "sphere":
            ue, ua = torch.rand((3, 200, 1), device="cuda"), torch.rand((3, 200, 1), device="cuda")
            ref = O.sphere_sampling(v.cpu(), q.cpu(), t.cpu(), ue.cpu(), ua.cpu())
        else:
            u = torch.rand((3, 200, 3), dtype=torch.float, device="cuda")
            ref = O.cuboid_sampling(v.cpu(), q.cpu(), t.cpu(), u.cpu())
        assert out.shape == (3, 200, 3)
        close(out, ref, atol=1e-6)
    assert ms.Sampling.cone_sampling(v, q, t, 10) is None
    with pytest.raises(AssertionError):
        ms.Sampling.sphere_sampling(v, q[:, :3], t, 10)


# ------------------------------------------------------------------------------------------------
# mesh surface sampling (TriangleMesh.sample, train_sphere.py:71-80)
# ------------------------------------------------------------------------------------------------
def _mesh386(golden_templates, b, seed):
    g = torch.Generator().manual_seed(seed)
    verts = torch.from_numpy(golden_templates["sphere386_vertices"]).float()
    faces = torch.from_numpy(golden_templates["sphere386_faces"]).long()
    offs = torch.tanh(torch.randn(b, verts.shape[0], 3, generator=g)) * 0.05        # SDNet-shaped vertex offsets
    return verts[None] + offs, faces, g


@pytest.mark.parametrize("b,n", [(1, 16384), (3, 1000), (2, 1)])
def test_mesh_sample_matches_oracle(vpn, O, golden_templates, b, n):
    verts, faces, g = _mesh386(golden_templates, b, 3)
    u = torch.rand(b, n, 3, generator=g)
    u[:, 0, 0] = 0.0                                   # first face
    if n > 2:
        u[:, 1, 0] = 0.99999994                        # largest draw below 1: last face with a non-zero share
        u[:, 2, 1] = 0.0                               # sqrt(0): the point is vertex 0 of its face
    vd = C(verts).requires_grad_()
    pts, fidx = vpn.sample_mesh_surface(vd, C(faces).int(), C(u))
    up = torch.rand(b, n, 3, generator=g)
    (pts * C(up)).sum().backward()
    for i in range(b):
        vo = verts[i].clone().requires_grad_()
        po, fo = O.mesh_sample(vo, faces, u[i])
        same(fidx[i].long(), fo, "face choice")                       # integer output: bit-exact
        same(pts[i], po.detach(), "points")                           # same fp32 operation order: bit-exact
        (po * up[i]).sum().backward()
        close(vd.grad[i], vo.grad, rtol=RTOL, atol=1e-5 * float(vo.grad.abs().max()), what="grad verts")


def test_mesh_sample_is_area_weighted(vpn, O, golden_templates):
    verts, faces, g = _mesh386(golden_templates, 1, 5)
    n = 400000
    u = torch.rand(1, n, 3, generator=g)
    pts, fidx = vpn.sample_mesh_surface(C(verts), C(faces).int(), C(u))
    share = np.diff(np.concatenate([[0.0], O.mesh_face_cdf(verts[0].numpy(), faces.numpy())]))
    freq = np.bincount(fidx[0].cpu().numpy(), minlength=faces.shape[0]) / n
    assert np.abs(freq - share).max() < 4 * np.sqrt(share.max() / n)       # 4 sigma of a binomial share
    # every point lies in the plane of, and inside, its triangle
    v = verts[0].numpy(); f = faces.numpy()[fidx[0].cpu().numpy()]
    p = pts[0].cpu().numpy()
    a, b_, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    nrm = np.cross(b_ - a, c - a); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    assert np.abs(((p - a) * nrm).sum(1)).max() < 1e-6


def test_dropin_triangle_mesh_sample(vpn, golden_templates):
    import modules.meshing as mm
    verts, faces, _ = _mesh386(golden_templates, 2, 9)
    meshes = [mm.TriangleMesh(C(verts[i]).requires_grad_(), C(faces)) for i in range(2)]
    pts, fidx = meshes[0].sample(4096)
    assert pts.shape == (4096, 3) and fidx.shape == (4096,) and fidx.dtype == torch.int64
    pts.sum().backward()
    close(meshes[0].vertices.grad.sum(), 3 * 4096.0, rtol=1e-5)
    batch = mm.TriangleMesh.sample_batch(meshes, 1024)
    assert batch.shape == (2, 1024, 3)
    r = batch.detach().norm(dim=2)
    assert float(r.min()) > 0.3 and float(r.max()) < 0.6           # 386.obj radius 0.45-0.47 + offsets <= 0.05*sqrt(3)


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs C3 / C5 at reduced batch: the whole step against the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kind,b,k,n,m,res,vertex", [
    # the oracle's Chamfer and rasteriser are dense ((P,M,3) and (H*W,F,6) temporaries): N, M and the render resolution
    # are reduced, K (hence the face count 16 128 / 32 256 and the primitive-major point layout) is the config's own;
    # full resolutions are covered by test_silhouette_vs_oracle / test_silhouette_loss_dropin with fewer faces
    ("c3", "cuboid", 1, 32, 1024, 2048, 32, False),      # C3: 32 cuboids x N samples + soft-silhouette L1
    ("c5", "sphere", 1, 128, 128, 4096, 32, True),       # C5 (train_gcn.py): 128 x 128 mesh vertices vs points + render
])
def test_step_config_shapes(vpn, O, name, kind, b, k, n, m, res, vertex):
    from vpn_b200 import templates
    g = torch.Generator().manual_seed(1234)
    v, q, t = O.synthetic_primitives(b, k)
    t = t * 0.35
    width = 2 if kind == "sphere" else 3
    u = torch.rand(b, k, n, width, generator=g)
    target = (torch.rand(b, m, 3, generator=g) - 0.5) * 0.9
    gt = (torch.rand(b, 1, res, res, generator=g) > 0.5).float()
    cfg = vpn.PrimitiveLossConfig(kind=kind, l_sil=1.0, vertex_chamfer=vertex)
    vc, qc, tc = (C(x).requires_grad_() for x in (v, q, t))
    out = vpn.PrimitiveLoss(cfg)(vc, qc, tc, None if vertex else C(u), C(target), silhouettes=C(gt))
    out["total"].backward()
    vo, qo, to = (x.clone().requires_grad_() for x in (v, q, t))
    tv, tf = templates.template(kind, "cpu")
    verts, faces = O.compose_primitive_meshes(tv, tf.long(), vo, qo, to)
    pts = verts if vertex else O.sample_predict_points(kind, vo, qo, to, u)
    ref = (O.chamfer_dense(pts, target) + 0.1 * O.chamfer_dense(to, target, w1=0.5, w2=1.0)
           + O.silhouette_loss(verts, faces, gt, torch.ones(b), torch.zeros(b), torch.zeros(b)))
    ref.backward()
    close(out["total"], ref, rtol=RTOL, what=f"{name} loss")
    # arg-mins of the big Chamfer: bit-exact against the oracle run on the device's own points
    gi = vpn.chamfer_nn(out["points"].detach(), C(target))
    ri = O.chamfer_nn(out["points"].detach().cpu(), target)
    same(gi[1].long(), ri[1], f"{name} idx1"); same(gi[3].long(), ri[3], f"{name} idx2")
    for got, want, nm in ((vc.grad, vo.grad, "v"), (qc.grad, qo.grad, "q"), (tc.grad, to.grad, "t")):
        close(got, want, rtol=1e-3, atol=2e-4 * float(want.abs().max()), what=f"{name} grad {nm}")


def test_step_with_canonical_frame_chamfer(vpn, O):
    """train.py:152-163 as written: view-frame Chamfer + Chamfer of view_to_obj_points(points) against the canonical
    targets + VP-diverse; eager step against the oracle, and the graph replay against the eager step (vertex mode: no
    random draw)."""
    g = torch.Generator().manual_seed(77)
    b, k, n, m = 2, 4, 512, 1024
    v, q, t = O.synthetic_primitives(b, k); t = t * 0.3
    u = torch.rand(b, k, n, 3, generator=g)
    tgt = (torch.rand(b, m, 3, generator=g) - 0.5) * 0.8
    dists = 1.0 + 0.5 * torch.rand(b, generator=g); elevs = 40.0 * torch.rand(b, generator=g)
    azims = 360.0 * torch.rand(b, generator=g); angles = 360.0 * torch.rand(b, generator=g)
    canon = tgt * dists[:, None, None]
    cfg = vpn.PrimitiveLossConfig(kind="cuboid", l_can_cd=1.0)
    vc, qc, tc = (C(x).requires_grad_() for x in (v, q, t))
    out = vpn.PrimitiveLoss(cfg)(vc, qc, tc, C(u), C(tgt), canonical_points=C(canon), dists=C(dists), elevs=C(elevs),
                                 azims=C(azims), angles=C(angles))
    out["total"].backward()
    vo, qo, to = (x.clone().requires_grad_() for x in (v, q, t))
    pts = O.sample_predict_points("cuboid", vo, qo, to, u)
    ref = (O.chamfer_dense(pts, tgt) + O.chamfer_dense(O.view_to_obj_points(pts, dists, elevs, azims, angles), canon)
           + 0.1 * O.chamfer_dense(to, tgt, w1=0.5, w2=1.0))
    ref.backward()
    close(out["total"], ref, rtol=RTOL, what="faithful loss")
    for got, want, nm in ((vc.grad, vo.grad, "v"), (qc.grad, qo.grad, "q"), (tc.grad, to.grad, "t")):
        close(got, want, rtol=1e-3, atol=2e-4 * float(want.abs().max()), what=f"faithful grad {nm}")
    cfgv = vpn.PrimitiveLossConfig(kind="sphere", l_can_cd=1.0, vertex_chamfer=True)
    cams = (C(dists), C(elevs), C(azims), C(angles))
    gr = vpn.GraphedPrimitiveLoss(cfgv, C(v), C(q), C(t), C(tgt), None, canonical_points=C(canon), cameras=cams)
    cams2 = (C(dists * 1.1), C(elevs + 3.0), C(azims), C(angles))
    for cam, cn in ((cams, canon), (cams2, canon * 0.9)):
        loss, gv, gq, gtt = gr(C(v), C(q), C(t), C(tgt), canonical_points=C(cn), cameras=cam)
        ve, qe, te = (C(x).requires_grad_() for x in (v, q, t))
        oe = vpn.PrimitiveLoss(cfgv)(ve, qe, te, None, C(tgt), canonical_points=C(cn), dists=cam[0], elevs=cam[1],
                                     azims=cam[2], angles=cam[3])
        oe["total"].backward()
        close(loss, oe["total"], rtol=1e-6)
        close(gv, ve.grad, rtol=1e-5, atol=1e-7); close(gq, qe.grad, rtol=1e-5, atol=1e-7); close(gtt, te.grad, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# EMD auction (modules/loss/emd): one launch per forward, deterministic; bit-exact against the restatement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b,n,eps,iters", [(3, 2048, 0.005, 50), (2, 1024, 0.002, 10), (2, 37, 0.005, 50), (1, 1, 0.005, 3),
                                            (2, 4096, 0.005, 20), (1, 5000, 0.01, 8)])
def test_emd_auction_matches_oracle(vpn, O, b, n, eps, iters):
    g = torch.Generator().manual_seed(100 + n)
    x1 = torch.rand(b, n, 3, generator=g)
    x2 = torch.rand(b, n, 3, generator=g)
    if n >= 64:
        x2[:, :8] = x2[:, 8:16]                    # duplicate objects: exact ties between best and second best
        x1[0, :4] = x1[0, 4:8]                     # duplicate bidders: equal increments on the same object
    xc = C(x1).requires_grad_()
    dist, ass = vpn.emd_auction(xc, C(x2), eps, iters)
    up = torch.rand(b, n, generator=g)
    (dist * C(up)).sum().backward()
    for i in range(b):
        d_ref, a_ref = O.emd_auction(x1[i].numpy(), x2[i].numpy(), eps, iters)
        same(ass[i], a_ref, f"assignment sample {i}")                  # integer output: bit-exact
        same(dist[i], d_ref, f"dist sample {i}")
        close(xc.grad[i], O.emd_backward(x1[i].numpy(), x2[i].numpy(), a_ref, up[i].numpy()), rtol=1e-6, atol=1e-7)


def test_emd_dropin_matches_reference_call(vpn):
    """train.py:188-195: dist, assignment = EarthMoverDistanceLoss()(predict_points, gt_points, 0.005, 50); sqrt(dist).mean()."""
    import modules.loss as ml
    g = torch.Generator().manual_seed(8)
    p = C(torch.rand(2, 2048, 3, generator=g)).requires_grad_()
    q = C(torch.rand(2, 2048, 3, generator=g))
    dist, ass = ml.EarthMoverDistanceLoss()(p, q, 0.005, 50)
    assert dist.shape == (2, 2048) and ass.shape == (2, 2048) and ass.dtype == torch.int32
    assert int(ass.min()) >= 0 and int(ass.max()) < 2048
    loss = torch.sqrt(dist).mean()
    loss.backward()
    assert torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0
    with pytest.raises(AssertionError):
        ml.EarthMoverDistanceLoss()(p[:, :1000], q[:, :1000], 0.005, 50)      # n % 1024 != 0 (emd_module.py:39)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_emd_auction_cluster_sizes_agree(vpn, O, cluster):
    """The CTAs-per-sample choice (thread-block cluster size) must not change the result."""
    from vpn_b200 import _lib
    lib = _lib.load()
    assert lib.vpn_set_tuning(b"emd_cluster", cluster) == 0
    try:
        g = torch.Generator().manual_seed(77)
        for n, iters in ((2048, 50), (600, 12), (13000, 3)):           # 13000: objects read through L2 instead of shared memory
            x1 = torch.rand(2, n, 3, generator=g); x2 = torch.rand(2, n, 3, generator=g)
            dist, ass = vpn.emd_auction(C(x1), C(x2), 0.005, iters)
            d_ref, a_ref = O.emd_auction(x1[1].numpy(), x2[1].numpy(), 0.005, iters)
            same(ass[1], a_ref, f"assignment n={n} cluster={cluster}")
            same(dist[1], d_ref, f"dist n={n} cluster={cluster}")
    finally:
        lib.vpn_set_tuning(b"emd_cluster", 0)


# ------------------------------------------------------------------------------------------------
# empty batches: every entry point returns empty outputs of the right shape and launches nothing harmful
# ------------------------------------------------------------------------------------------------
def test_empty_batch(vpn, golden_templates):
    dev = "cuda"
    z = lambda *s: torch.zeros(*s, device=dev)
    pts = vpn.sample_primitives("cuboid", z(0, 4, 3), z(0, 4, 4), z(0, 4, 3), z(0, 4, 16, 3))
    assert pts.shape == (0, 64, 3)
    m1, i1, m2, i2 = vpn.chamfer_nn(z(0, 2048, 3), z(0, 256, 3))
    assert m1.shape == (0, 2048) and i2.shape == (0, 256) and i1.dtype == torch.int32
    assert vpn.chamfer_distance(z(0, 2048, 3), z(0, 256, 3), each_batch=True).shape == (0,)
    tv = torch.from_numpy(golden_templates["sphere_vertices"]).float().to(dev)
    tf = torch.from_numpy(golden_templates["sphere_faces"]).int().to(dev)
    verts = vpn.mesh_vertices(tv, z(0, 2, 3), z(0, 2, 4), z(0, 2, 3))
    assert verts.shape == (0, 2 * tv.shape[0], 3)
    alpha, covered, _ = vpn.soft_silhouette(verts, torch.cat([tf, tf + tv.shape[0]]), z(0, 3, 3), z(0, 3), 32, 32)
    assert alpha.shape == (0, 32, 32) and covered.shape == (0, 32, 32)
    p, f = vpn.sample_mesh_surface(z(0, tv.shape[0], 3), tf, z(0, 100, 3))
    assert p.shape == (0, 100, 3) and f.shape == (0, 100)
    d, a = vpn.emd_auction(z(0, 1024, 3), z(0, 1024, 3), 0.005, 10)
    assert d.shape == (0, 1024) and a.shape == (0, 1024)
    torch.cuda.synchronize()


def test_graphed_step_matches_eager(vpn, O):
    """GraphedPrimitiveLoss replays the captured step: same loss / gradients as the eager step on the same uniforms is not
    testable (the graph draws its own), so compare a vertex-Chamfer step (no random draw) and check the sampled one runs."""
    g = torch.Generator().manual_seed(3)
    b, k, m, res = 2, 8, 1024, 32
    v, q, t = O.synthetic_primitives(b, k); t = t * 0.3
    tgt = (torch.rand(b, m, 3, generator=g) - 0.5) * 0.8
    gt = (torch.rand(b, 1, res, res, generator=g) > 0.5).float()
    cfg = vpn.PrimitiveLossConfig(kind="sphere", l_sil=1.0, vertex_chamfer=True)
    gr = vpn.GraphedPrimitiveLoss(cfg, C(v), C(q), C(t), C(tgt), C(gt))
    for scale in (1.0, 0.7):                                   # second call: new inputs through the static buffers
        vc, qc, tc = (C(x * scale if i == 0 else x).requires_grad_() for i, x in enumerate((v, q, t)))
        loss, gv, gq, gtt = gr(vc.detach(), qc.detach(), tc.detach(), C(tgt), C(gt))
        out = vpn.PrimitiveLoss(cfg)(vc, qc, tc, None, C(tgt), silhouettes=C(gt))
        out["total"].backward()
        close(loss, out["total"], rtol=1e-6)
        close(gv, vc.grad, rtol=1e-5, atol=1e-7); close(gq, qc.grad, rtol=1e-5, atol=1e-7); close(gtt, tc.grad, rtol=1e-5, atol=1e-7)
    assert gr.launches_per_step > 5
    cfg2 = vpn.PrimitiveLossConfig(kind="cuboid")
    gs = vpn.GraphedPrimitiveLoss(cfg2, C(v), C(q), C(t), C(tgt), None, n_samples=256)
    l1 = gs(C(v), C(q), C(t), C(tgt))[0].item()
    l2 = gs(C(v), C(q), C(t), C(tgt))[0].item()
    assert l1 > 0 and l2 > 0 and l1 != l2                      # a fresh uniform draw every replay


def test_host_pipeline_matches_direct_steps(vpn, O):
    """vpn_b200.HostPipeline (pinned host batches, copies overlapped with the previous step on a copy stream): the same
    losses and gradients (to the repeatability of the atomics in the backward), in order, as calling the graphed step on device tensors (vertex Chamfer + silhouette: no random
    draw); six batches through the two staging slots."""
    g = torch.Generator().manual_seed(11)
    b, k, m, res = 2, 8, 1024, 32
    v, q, t = O.synthetic_primitives(b, k); t = t * 0.3
    cfg = vpn.PrimitiveLossConfig(kind="sphere", l_sil=1.0, vertex_chamfer=True)
    batches = []
    for i in range(6):
        batches.append({"v": (v * (1.0 - 0.05 * i)).contiguous().pin_memory(), "q": q.clone().pin_memory(), "t": (t + 0.01 * i).pin_memory(),
                        "target": ((torch.rand(b, m, 3, generator=g) - 0.5) * 0.8).pin_memory(),
                        "sil": (torch.rand(b, 1, res, res, generator=g) > 0.5).float().pin_memory(), "unused": None})
    d0 = batches[0]
    gr = vpn.GraphedPrimitiveLoss(cfg, C(d0["v"]), C(d0["q"]), C(d0["t"]), C(d0["target"]), C(d0["sil"]))
    want = []
    for bt in batches:
        outs = gr(C(bt["v"]), C(bt["q"]), C(bt["t"]), C(bt["target"]), C(bt["sil"]))
        want.append([o.detach().cpu().clone() for o in outs])
    pipe = vpn.HostPipeline(lambda s: gr(s["v"], s["q"], s["t"], s["target"], s["sil"]), d0, "cuda")
    got = []
    pipe.submit(batches[0])
    for bt in batches[1:]:
        pipe.submit(bt)
        got.append([o.clone() for o in pipe.result()])
    got.append([o.clone() for o in pipe.result()])
    for w, gg in zip(want, got):                                 # the backward scatters with atomics: not bitwise repeatable
        close(gg[0], w[0], rtol=1e-6)
        for a, c in zip(w[1:], gg[1:]):
            close(c, a, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------
# GCN vertex-feature pooling (modules/network/gcn.py:84-164)
# ------------------------------------------------------------------------------------------------
def test_feature_pooling_golden(vpn, golden):
    """Against what the reference's own GCNModel.get_bound_of_images / perceptual_feature_pooling produced."""
    g = golden
    imgs, pts = C(g["in_pool_imgs"]), C(g["in_pool_points"])
    feats = [C(g[f"in_pool_feat{i}"]).requires_grad_() for i in range(3)]
    bounds = vpn.image_bounds(imgs)
    same(bounds, g["ref_pool_bounds"], "bounds")                      # integer pixel indices, same normalisation ops
    pg = pts.clone().requires_grad_()
    out = vpn.perceptual_feature_pooling(feats, pg, bounds)
    close(out, g["ref_pool_out"], rtol=RTOL, atol=1e-6, what="pooled features")
    (out * C(g["in_pool_upstream"])).sum().backward()
    for i, f in enumerate(feats):
        close(f.grad, g[f"ref_pool_grad_feat{i}"], rtol=RTOL, atol=1e-5, what=f"grad feat {i}")
    ref = g["ref_pool_grad_points"]
    close(pg.grad, ref, rtol=1e-3, atol=1e-4 * float(np.abs(ref).max()), what="grad points")
    import modules
    close(modules.GCNFeaturePooling.get_local_features(pts, imgs, [f.detach() for f in feats]), g["ref_pool_local"],
          rtol=RTOL, atol=1e-6, what="get_local_features")


@pytest.mark.parametrize("b,n,maps", [
    (4, 2048, [(64, 35, 35), (128, 18, 18), (256, 9, 9), (512, 5, 5)]),      # train_gcn.py: ResNet-18 maps of a 137 x 137 view
    (2, 300, [(3, 7, 5), (33, 6, 6), (1, 1, 1)]),                             # ragged channel counts, 1 x 1 map
    (1, 37, [(2, 200, 200)]),                                                 # plane larger than the shared-memory budget
])
def test_feature_pooling_vs_oracle(vpn, O, b, n, maps):
    g = torch.Generator().manual_seed(11)
    pts = (torch.rand(b, n, 3, generator=g) - 0.5) * 0.9
    imgs = torch.zeros(b, 3, 137, 137)
    for i in range(b):
        x0, y0 = 10 + 7 * i, 20 + 5 * i
        imgs[i, :, y0:y0 + 60, x0:x0 + 80] = torch.rand(3, 60, 80, generator=g)
    feats = [torch.randn(b, c, h, w, generator=g) for c, h, w in maps]
    bounds_ref = O.image_bounds(imgs)
    bounds = vpn.image_bounds(C(imgs))
    same(bounds, bounds_ref, "bounds")
    fo = [f.clone().requires_grad_() for f in feats]
    po = pts.clone().requires_grad_()
    ref = O.perceptual_feature_pooling(fo, po, bounds_ref)
    up = torch.randn(ref.shape, generator=g)
    (ref * up).sum().backward()
    fc = [C(f).requires_grad_() for f in feats]
    pc = C(pts).requires_grad_()
    out = vpn.perceptual_feature_pooling(fc, pc, bounds)
    assert out.shape == (b, n, sum(c for c, _, _ in maps)) and out.is_contiguous()
    close(out, ref, rtol=RTOL, atol=2e-6, what="pooled")
    (out * C(up)).sum().backward()
    for i in range(len(maps)):
        close(fc[i].grad, fo[i].grad, rtol=RTOL, atol=1e-5 * max(1.0, float(fo[i].grad.abs().max())), what=f"grad feat {i}")
    close(pc.grad, po.grad, rtol=1e-3, atol=2e-4 * float(po.grad.abs().max()), what="grad points")


def test_feature_pooling_backward_edge_geometry(vpn, O):
    """The cell-sorted backward on the geometries that stress its cell / tap-validity logic: vertices exactly on texel
    centres (integer pixel coordinates: a tap's coefficients vanish for that vertex while the tap is valid for its cell
    mates), image bounds beyond [-1, 1] (taps and whole cells outside the plane), all vertices in one cell, a 1-row
    map, and the two backward implementations (shared-memory atomics / cell-sorted) against each other and the oracle."""
    import ctypes
    from vpn_b200 import _lib
    from vpn_b200._lib import ptr, stream_ptr, check
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    b, n = 2, 700
    maps = [(5, 9, 11), (40, 5, 5), (3, 1, 6), (70, 4, 3)]
    for bounds in (torch.tensor([[-1.0, 1.0, -1.0, 1.0], [-1.0, 1.0, -1.0, 1.0]]),          # exact texel hits with the lattice below
                   torch.tensor([[-1.6, 1.3, -0.2, 2.0], [0.5, 0.6, -3.0, -1.2]])):         # partly / wholly outside the plane
        pts = torch.rand(b, n, 3, generator=g)
        pts[0, :400, 1:] = torch.round(pts[0, :400, 1:] * 10) / 10        # lattice: with bounds +-1 and a 11-wide map -> integer ix
        pts[1, :, 1:] = 0.5 + 0.01 * torch.rand(n, 2, generator=g)        # everything in one or two cells ...
        pts[1, 0, 1:] = 0.0; pts[1, 1, 1:] = 1.0                          # ... except the two vertices that set the range
        feats = [torch.randn(b, c, h, w, generator=g) for c, h, w in maps]
        fo = [f.clone().requires_grad_() for f in feats]
        po = pts.clone().requires_grad_()
        ref = O.perceptual_feature_pooling(fo, po, bounds)
        up = torch.randn(ref.shape, generator=g)
        (ref * up).sum().backward()
        fc = [C(f).requires_grad_() for f in feats]
        pc = C(pts).requires_grad_()
        out = vpn.perceptual_feature_pooling(fc, pc, C(bounds))
        close(out, ref, rtol=RTOL, atol=2e-6, what="pooled")
        (out * C(up)).sum().backward()
        for i in range(len(maps)):
            close(fc[i].grad, fo[i].grad, rtol=RTOL, atol=1e-5 * max(1.0, float(fo[i].grad.abs().max())), what=f"grad feat {i}")
        close(pc.grad, po.grad, rtol=1e-3, atol=2e-4 * float(po.grad.abs().max()), what="grad points")
        # old kernel == new kernels, map by map, through the C ABI
        rng = torch.empty(b, 4, device="cuda"); arg = torch.empty(b, 4, dtype=torch.int32, device="cuda")
        check(lib.vpn_points_yz_range(ptr(pc.detach()), ptr(rng), ptr(arg), b, n, stream_ptr("cuda")), "range")
        nws = ctypes.c_size_t(0)
        check(lib.vpn_feature_pool_bwd_workspace_bytes(b, n, ctypes.byref(nws)), "ws")
        ws = torch.empty(nws.value, dtype=torch.uint8, device="cuda")
        ctot, coff = sum(c for c, _, _ in maps), 0
        gout = C(up).contiguous()
        bd = C(bounds).contiguous()
        for (c, h, w), f in zip(maps, feats):
            fcu = C(f)
            g_old, g_new = torch.empty_like(fcu), torch.empty_like(fcu)
            gg_old, gg_new = torch.zeros(b, n, 2, device="cuda"), torch.zeros(b, n, 2, device="cuda")
            check(lib.vpn_feature_pool_bwd(ptr(fcu), ptr(pc.detach()), ptr(bd), ptr(rng), ptr(gout), ptr(g_old), ptr(gg_old), b, c, h, w, n,
                                           ctot, coff, stream_ptr("cuda")), "old")
            check(lib.vpn_feature_pool_bwd_sorted(ptr(fcu), ptr(pc.detach()), ptr(bd), ptr(rng), ptr(gout), ptr(g_new), ptr(gg_new), ptr(ws),
                                                  nws.value, b, c, h, w, n, ctot, coff, stream_ptr("cuda")), "new")
            close(g_new, g_old, rtol=RTOL, atol=1e-5 * max(1.0, float(g_old.abs().max())), what=f"map {c}x{h}x{w}: grad_feat old vs sorted")
            close(gg_new, gg_old, rtol=1e-3, atol=1e-4 * max(1.0, float(gg_old.abs().max())), what=f"map {c}x{h}x{w}: grad_grid old vs sorted")
            coff += c


def test_image_bounds_edge_cases(vpn, O):
    g = torch.Generator().manual_seed(5)
    cases = [torch.zeros(2, 3, 16, 12)]                                        # empty masks: bounds stay (0, w, 0, h)
    im = torch.zeros(3, 4, 9, 31)
    im[0, :, :, 0] = 1.0; im[0, :, 4, 17] = 1.0                               # occupied column 0 is skipped as a lower bound
    im[1, 0] = 0.0299; im[1, 1, 8, 30] = 0.001                                 # just below / above the 0.03 threshold
    im[2] = torch.rand(4, 9, 31, generator=g)                                  # everything occupied
    cases.append(im)
    cases.append(torch.rand(5, 1, 137, 137, generator=g) * (torch.rand(5, 1, 137, 137, generator=g) > 0.999))
    for imgs in cases:
        same(vpn.image_bounds(C(imgs)), O.image_bounds(imgs), "bounds")
    import modules
    with pytest.raises(AssertionError):
        modules.GCNFeaturePooling.get_bound_of_images(C(torch.zeros(3, 8, 8)))
    with pytest.raises(AssertionError):
        modules.GCNFeaturePooling.perceptual_feature_pooling([C(torch.zeros(1, 2, 4, 4))], C(torch.zeros(5, 3)), C(torch.zeros(1, 4)))
