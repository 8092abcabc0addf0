"""Round-2 GPU parity tests the round-1 verdict asked for: the rasteriser at BASELINE sizes (C3: F = 16 128 at 128x128,
C5: F = 32 256 at 256x256) on pixel subsets, non-default cameras, both readings of the soft-pass back-face rule, MSE,
the canonical-frame Chamfer together with the silhouette (eager and graphed), obj_to_view_points backward, and the
tensor-core Chamfer filter directly against arg-mins the REFERENCE produced (tests/golden/round2_golden.npz).
Render parity stays "matches our DIB-R restatement" (Kaolin v0.1 is not installable)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-4


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


def _scene(vpn, O, kind, b, k, seed, spread=0.3, scale=1.5):
    from vpn_b200 import templates
    v, q, t = O.synthetic_primitives(b, k, seed=seed)
    tv, tf = templates.template(kind, "cpu")
    verts = vpn.mesh_vertices(tv.cuda(), (v * scale).cuda(), q.cuda(), (t * spread).cuda()).detach()
    faces = torch.cat([tf + i * tv.shape[0] for i in range(k)])
    return verts, faces


def _cams(O, dists, elevs, azims):
    cams = [O.look_at_camera(float(a), float(e), float(d)) for d, e, a in zip(dists, elevs, azims)]
    return torch.stack([c[0] for c in cams]), torch.stack([c[1] for c in cams])


def _compare(vpn, O, verts, faces, dists, elevs, azims, res, cull, pixels=None, wseed=4):
    """alpha and vertex gradients, CUDA vs oracle, on the device's own vertices.  pixels: flat indices or None (all)."""
    b = verts.shape[0]
    rot, pos = _cams(O, dists, elevs, azims)
    r_c, p_c = vpn.look_at_cameras(C(azims).float(), C(elevs).float(), C(dists).float())
    close(r_c, rot, atol=2e-7, what="camera rotation"); close(p_c, pos, atol=2e-7, what="camera position")
    vg = verts.clone().requires_grad_()
    alpha, covered, _ = vpn.soft_silhouette(vg, faces.cuda().int(), r_c, p_c, res, res, soft_cull_backfaces=cull)
    vr = verts.detach().cpu().clone().requires_grad_()
    # cameras: the oracle gets the device's fp32 camera so that only the rasteriser's arithmetic is compared
    ref = O.soft_silhouette(vr, faces.long(), r_c.cpu(), p_c.cpu(), res, res, soft_cull_backfaces=cull, pixels=pixels)
    gen = torch.Generator().manual_seed(wseed)
    if pixels is None:
        got = alpha
        w = torch.rand(b, res, res, generator=gen)
        wg = w.cuda()
    else:
        got = alpha.reshape(b, -1)[:, pixels.cuda()]
        w = torch.rand(b, pixels.numel(), generator=gen)
        wg = torch.zeros(b, res * res, device="cuda")
        wg[:, pixels.cuda()] = w.cuda()
        wg = wg.view(b, res, res)
    close(got, ref, rtol=RTOL, atol=2e-6, what="alpha")
    assert ((got.detach().cpu() == 1.0) == (ref.detach() == 1.0)).all(), "coverage differs"
    (ref * w).sum().backward()
    (alpha * wg).sum().backward()
    gref = vr.grad
    close(vg.grad, gref, rtol=1e-3, atol=1e-4 * float(gref.abs().max()), what="vertex gradient")
    return alpha.detach()


@pytest.mark.parametrize("cull", [False, True])
@pytest.mark.parametrize("kind,b,k,res", [("sphere", 2, 3, 32), ("cuboid", 2, 2, 48), ("sphere", 1, 16, 64)])
def test_silhouette_both_backface_rules(vpn, O, kind, b, k, res, cull):
    """soft_cull_backfaces False (DIB-R: back faces culled by the coverage pass only) and True, against the restatement."""
    verts, faces = _scene(vpn, O, kind, b, k, seed=21)
    one, zero = torch.ones(b), torch.zeros(b)
    a = _compare(vpn, O, verts, faces, one, zero, zero, res, cull)
    if kind == "sphere" and not cull:
        ac, _, _ = vpn.soft_silhouette(verts, faces.cuda().int(), *vpn.look_at_cameras(zero.cuda(), zero.cuda(), one.cuda()),
                                       res, res, soft_cull_backfaces=True)
        assert (a >= ac - 1e-6).all() and (a > ac + 1e-5).any(), "back faces must add soft coverage on a closed mesh"


def test_silhouette_inward_wound_mesh(vpn, O, golden_templates):
    """386.obj is wound inward: from outside, the far hemisphere faces the camera.  Both rules."""
    tv = torch.tensor(golden_templates["sphere386_vertices"]); tf = torch.tensor(golden_templates["sphere386_faces"])
    gen = torch.Generator().manual_seed(5)
    verts = (tv[None] + 0.02 * torch.randn(2, 386, 3, generator=gen)).cuda()
    for cull in (False, True):
        _compare(vpn, O, verts, tf, torch.ones(2), torch.zeros(2), torch.zeros(2), 64, cull)


@pytest.mark.parametrize("cam", [(1.0, 20.0, 90.0), (1.3, 25.0, 200.0), (1.5, 40.0, 359.0), (1.2, -15.0, 45.0), (2.0, 0.0, 180.0)])
def test_silhouette_cameras(vpn, O, cam):
    """Non-default cameras: the dist / elev / azim ranges bench.py's faithful workload and dataset.py use."""
    b = 2
    verts, faces = _scene(vpn, O, "cuboid", b, 4, seed=int(cam[2]))
    d = torch.tensor([cam[0], cam[0] * 1.1]); e = torch.tensor([cam[1], cam[1] * 0.5]); a = torch.tensor([cam[2], cam[2] + 33.0])
    _compare(vpn, O, verts, faces, d, e, a, 48, False)


def _pixel_subset(alpha, n_soft=320, n_other=96, seed=0):
    """Flat pixel indices: every kind of pixel, weighted towards the soft rim (0 < alpha < 1), where the order-dependent
    first-knum rule and the distance arithmetic matter."""
    a = alpha.reshape(alpha.shape[0], -1).cpu()
    gen = torch.Generator().manual_seed(seed)
    soft = ((a > 0) & (a < 1)).any(0).nonzero().flatten()
    cov = (a == 1).any(0).nonzero().flatten()
    emp = (a == 0).all(0).nonzero().flatten()
    pick = lambda s, n: s[torch.randperm(s.numel(), generator=gen)[:n]]
    return torch.unique(torch.cat([pick(soft, n_soft), pick(cov, n_other), pick(emp, n_other)]))


@pytest.mark.parametrize("name,kind,b,k,res,cam", [
    ("c3", "cuboid", 2, 32, 128, (1.0, 0.0, 0.0)),          # BASELINE configs[2]: 32 cuboids -> F = 16 128, 128 x 128
    ("c3-sphere", "sphere", 1, 32, 128, (1.0, 0.0, 0.0)),   # same with sphere templates: F = 8 064
    ("c5", "sphere", 1, 128, 256, (1.0, 0.0, 0.0)),         # BASELINE configs[4]: 128 spheres -> F = 32 256, 256 x 256
    ("c5-cam", "sphere", 1, 128, 256, (1.4, 30.0, 120.0)),
])
@pytest.mark.parametrize("cull", [False, True])
def test_silhouette_baseline_sizes(vpn, O, name, kind, b, k, res, cam, cull):
    """Full face counts and resolutions of BASELINE C3 / C5; the dense oracle evaluates a ~500-pixel subset chosen from
    the rendered image (soft-rim pixels first), alpha AND vertex gradients (upstream gradient non-zero on the subset)."""
    verts, faces = _scene(vpn, O, kind, b, k, seed=1234, spread=0.35, scale=1.0)
    assert faces.shape[0] == {"cuboid": 504, "sphere": 252}[kind] * k
    d, e, a = (torch.full((b,), float(x)) for x in cam)
    r_c, p_c = vpn.look_at_cameras(a.cuda(), e.cuda(), d.cuda())
    alpha, _, _ = vpn.soft_silhouette(verts, faces.cuda().int(), r_c, p_c, res, res, soft_cull_backfaces=cull)
    pix = _pixel_subset(alpha)
    assert pix.numel() >= 300
    _compare(vpn, O, verts, faces, d, e, a, res, cull, pixels=pix)


@pytest.mark.parametrize("loss_name", ["L1", "MSE"])
def test_silhouette_loss_l1_and_mse(vpn, O, loss_name):
    """silhouette.py:11: L1Loss or MSELoss by SILHOUETTE_LOSS_FUNC, through the batched step."""
    from vpn_b200 import templates
    b, k, res = 2, 4, 64
    v, q, t = O.synthetic_primitives(b, k, seed=9); t = t * 0.25; v = v * 1.5
    g = torch.Generator().manual_seed(2)
    tgt = (torch.rand(b, 256, 3, generator=g) - 0.5) * 0.8
    gt = (torch.rand(b, 1, res, res, generator=g) > 0.5).float()
    cfg = vpn.PrimitiveLossConfig(kind="sphere", l_sil=1.0, l_view_cd=0.0, l_vp_div=0.0, silhouette_loss=loss_name, vertex_chamfer=True)
    vc, qc, tc = (C(x).requires_grad_() for x in (v, q, t))
    out = vpn.PrimitiveLoss(cfg)(vc, qc, tc, None, C(tgt), silhouettes=C(gt))
    out["total"].backward()
    tv, tf = templates.template("sphere", "cpu")
    vo, qo, to = (x.clone().requires_grad_() for x in (v, q, t))
    verts, faces = O.compose_primitive_meshes(tv, tf.long(), vo, qo, to)
    ref = O.silhouette_loss(verts, faces, gt, torch.ones(b), torch.zeros(b), torch.zeros(b), loss_func=loss_name)
    ref.backward()
    close(out["total"], ref, rtol=RTOL, atol=1e-7)
    for got, want, nm in ((vc.grad, vo.grad, "v"), (qc.grad, qo.grad, "q"), (tc.grad, to.grad, "t")):
        close(got, want, rtol=2e-3, atol=3e-4 * float(want.abs().max()), what=f"{loss_name} grad {nm}")


def test_step_canonical_chamfer_with_silhouette(vpn, O):
    """train.py:152-176 with L_CAN_CD != 0 AND L_SIL != 0: the view parameters feed view_to_obj_points only; the
    silhouette is rendered from the view-centred camera (dist 1, elev 0, azim 0), eager and from the CUDA graph."""
    from vpn_b200 import templates
    g = torch.Generator().manual_seed(31)
    b, k, m, res = 2, 4, 512, 32
    v, q, t = O.synthetic_primitives(b, k, seed=5); t = t * 0.3; v = v * 1.5
    tgt = (torch.rand(b, m, 3, generator=g) - 0.5) * 0.8
    gt = (torch.rand(b, 1, res, res, generator=g) > 0.5).float()
    dists = 1.0 + 0.5 * torch.rand(b, generator=g); elevs = 20 + 20.0 * torch.rand(b, generator=g)
    azims = 360.0 * torch.rand(b, generator=g); angles = 360.0 * torch.rand(b, generator=g)
    canon = tgt * dists[:, None, None]
    cfg = vpn.PrimitiveLossConfig(kind="sphere", l_can_cd=1.0, l_sil=1.0, vertex_chamfer=True)
    vc, qc, tc = (C(x).requires_grad_() for x in (v, q, t))
    out = vpn.PrimitiveLoss(cfg)(vc, qc, tc, None, C(tgt), silhouettes=C(gt), canonical_points=C(canon), dists=C(dists),
                                 elevs=C(elevs), azims=C(azims), angles=C(angles))
    out["total"].backward()
    tv, tf = templates.template("sphere", "cpu")
    vo, qo, to = (x.clone().requires_grad_() for x in (v, q, t))
    verts, faces = O.compose_primitive_meshes(tv, tf.long(), vo, qo, to)
    ref = (O.chamfer_dense(verts, tgt) + O.chamfer_dense(O.view_to_obj_points(verts, dists, elevs, azims, angles), canon)
           + 0.1 * O.chamfer_dense(to, tgt, w1=0.5, w2=1.0)
           + O.silhouette_loss(verts, faces, gt, torch.ones(b), torch.zeros(b), torch.zeros(b)))
    ref.backward()
    close(out["total"], ref, rtol=RTOL, what="loss with canonical Chamfer + silhouette")
    for got, want, nm in ((vc.grad, vo.grad, "v"), (qc.grad, qo.grad, "q"), (tc.grad, to.grad, "t")):
        close(got, want, rtol=2e-3, atol=3e-4 * float(want.abs().max()), what=f"grad {nm}")
    cams = (C(dists), C(elevs), C(azims), C(angles))
    gr = vpn.GraphedPrimitiveLoss(cfg, C(v), C(q), C(t), C(tgt), C(gt), canonical_points=C(canon), cameras=cams)
    loss, gv, gq, gtt = gr(C(v), C(q), C(t), C(tgt), C(gt), canonical_points=C(canon), cameras=cams)
    close(loss, out["total"], rtol=1e-6)
    close(gv, vc.grad, rtol=1e-5, atol=1e-7); close(gq, qc.grad, rtol=1e-5, atol=1e-7); close(gtt, tc.grad, rtol=1e-5, atol=1e-7)


def test_view_transform_backward_golden(vpn, golden, golden2):
    """obj_to_view_points / rotate_points_forward_x_axis backward against the reference's autograd (transform.py:50-94)."""
    import modules.transform as mt
    g = golden
    d, e, a, ang = (C(g["in_tf_" + k]) for k in ("dists", "elevs", "azims", "angles"))
    w = C(g["in_tf_upstream"])
    pg = C(g["in_tf_points"]).requires_grad_()
    (mt.obj_to_view_points(pg, d, e, a) * w).sum().backward()
    close(pg.grad, golden2["ref_tf_obj_to_view_grad_points"], atol=3e-6)
    pg = C(g["in_tf_points"]).requires_grad_()
    (mt.rotate_points_forward_x_axis(pg, ang) * w).sum().backward()
    close(pg.grad, golden2["ref_tf_rotate_x_grad_points"], atol=2e-6)


@pytest.mark.parametrize("impl", [0, 5, 4, 2, 1])
def test_chamfer_reference_argmins_every_impl(vpn, golden2, impl):
    """Arg-mins produced by the REFERENCE's torch.min on P = 1536, M = 640 clouds, checked directly against the
    tensor-core filter (impl 5), auto (0), the CUDA-core filters and the generic kernel."""
    g = golden2
    p1, p2 = C(g["in_cd_p1"]), C(g["in_cd_p2"])
    name = vpn.chamfer_main_kernel_name(p1.shape[0], p1.shape[1], p2.shape[1], impl)
    if impl in (0, 5):
        assert name == "chamfer_tc_kernel", name
    m1, i1, m2, i2 = vpn.chamfer_nn(p1, p2, impl)
    same(i1, g["ref_cd_idx1"], f"idx1 impl {impl}"); same(i2, g["ref_cd_idx2"], f"idx2 impl {impl}")
    close(m1, g["ref_cd_min1"], rtol=1.3e-7, atol=0); close(m2, g["ref_cd_min2"], rtol=1.3e-7, atol=0)
    a, b = p1.clone().requires_grad_(), p2.clone().requires_grad_()
    loss = vpn.chamfer_distance(a, b, impl=impl)
    close(loss, g["ref_cd_loss"], rtol=1e-6, atol=0)
    loss.backward()
    close(a.grad, g["ref_cd_grad_p1"], rtol=1e-4, atol=1e-9); close(b.grad, g["ref_cd_grad_p2"], rtol=1e-4, atol=1e-9)
    close(vpn.chamfer_distance(p1, p2, each_batch=True, impl=impl), g["ref_cd_loss_each"], rtol=1e-6, atol=0)
