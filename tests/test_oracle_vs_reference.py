"""Live cross-check: the oracle against the REFERENCE's own modules on fresh random inputs, several seeds.

Runs only where the reference tree exists (the build container: /root/reference); on the GPU box it is skipped and
the committed golden vectors (tests/test_oracle_golden.py) carry the pin.  Same accommodations as
oracle/make_golden.py (oracle/ref_import.py documents them)."""
import numpy as np
import pytest
import torch

from oracle import ref_import
from oracle import vpn_oracle as O

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def R():
    return ref_import.load_reference()


def eq(a, b, what=""):
    np.testing.assert_array_equal(a.detach().numpy(), b.detach().numpy(), err_msg=what)


def close(a, b, rtol=1e-5, atol=1e-6, what=""):
    np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=rtol, atol=atol, err_msg=what)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_transform_and_sampling_bit_exact(R, seed):
    g = torch.Generator().manual_seed(100 + seed)
    B, n = 3, 33
    pts = torch.randn(B, n, 3, generator=g)
    q = torch.randn(B, 4, generator=g) * 1.5                       # turn fractions outside [0, 1) exercise the modulo
    t = torch.randn(B, 3, generator=g)
    eq(O.transform_points(pts, q, t), R.transform.transform_points(pts, q, t), "transform_points")
    d, e = torch.rand(B, generator=g) + 0.5, torch.rand(B, generator=g) * 80 - 20
    a, ang = torch.rand(B, generator=g) * 360, torch.rand(B, generator=g) * 360
    eq(O.view_to_obj_points(pts, d, e, a, ang), R.transform.view_to_obj_points(pts, d, e, a, ang), "view_to_obj")
    eq(O.obj_to_view_points(pts, d, e, a), R.transform.obj_to_view_points(pts, d, e, a), "obj_to_view")
    eq(O.rotate_points_forward_x_axis(pts, ang), R.transform.rotate_points_forward_x_axis(pts, ang), "rotate_x")
    N = 50 + 7 * seed
    v = (torch.sigmoid(torch.randn(B, 3, generator=g)) + 0.1) / torch.tensor([8.0, 10.0, 10.0])
    qs, ts = torch.sigmoid(torch.randn(B, 4, generator=g)), torch.tanh(torch.randn(B, 3, generator=g))
    ue, ua = torch.rand(B, N, 1, generator=g), torch.rand(B, N, 1, generator=g)
    with ref_import.forced_uniforms([ue, ua]):
        ref = R.Sampling.sphere_sampling(v, qs, ts, N)
    eq(O.sphere_sampling(v, qs, ts, ue, ua), ref, "sphere_sampling")
    u = torch.rand(B, N, 3, generator=g)
    with ref_import.forced_uniforms([u]):
        ref = R.Sampling.cuboid_sampling(v, qs, ts, N)
    eq(O.cuboid_sampling(v, qs, ts, u), ref, "cuboid_sampling")
    eq(O.cuboid_face_counts(v, 4096), R.cuboid.get_faces_points(v[:, 0:1], v[:, 1:2], v[:, 2:3], 4096), "face counts")


@pytest.mark.parametrize("seed", [0, 1])
def test_chamfer_and_vp_diverse(R, seed):
    g = torch.Generator().manual_seed(200 + seed)
    B, P, M = 2, 90, 41
    p1 = (torch.randn(B, P, 3, generator=g) * 0.3).requires_grad_()
    p2 = (torch.randn(B, M, 3, generator=g) * 0.3).requires_grad_()
    cd = R.ChamferDistanceLoss()
    ref = cd(p1, p2)
    gr = torch.autograd.grad(ref, (p1, p2))
    got = O.chamfer_dense(p1, p2)
    go = torch.autograd.grad(got, (p1, p2))
    eq(got, ref, "chamfer loss")
    # gradients: the same autograd graph, but torch-CPU's sqrt (MKL VML) is 1 ulp off on some inputs depending on how a
    # buffer is vectorised, which moves a few gradient entries by ~1e-9
    close(go[0], gr[0], rtol=1e-4, atol=1e-8, what="grad p1"); close(go[1], gr[1], rtol=1e-4, atol=1e-8, what="grad p2")
    eq(O.chamfer_dense(p1, p2, each_batch=True, w1=0.5, w2=2.0), cd(p1, p2, each_batch=True, w1=0.5, w2=2.0), "each_batch")
    # the O(P+M)-memory search the GPU tests use as their arg-min authority: same winners as torch.min on the dense matrix
    diff = p1[:, :, None, :] - p2[:, None, :, :]
    dist = torch.sum(diff * diff, dim=3)
    i1 = torch.min(torch.sqrt(dist), dim=2)[1]
    i2 = torch.min(torch.sqrt(dist.transpose(1, 2)), dim=2)[1]
    m1, j1, m2, j2 = O.chamfer_nn(p1.detach(), p2.detach())
    eq(j1, i1, "idx1"); eq(j2, i2, "idx2")
    K = R.config.VP_NUM
    tr = [torch.tanh(torch.randn(B, 3, generator=g)) for _ in range(K)]
    eq(O.vp_diverse(tr, p2.detach()), R.VPDiverseLoss()(tr, p2.detach()), "vp_diverse")


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_gcn_feature_pooling(R, seed):
    g = torch.Generator().manual_seed(300 + seed)
    B, N = 2, 57
    imgs = torch.zeros(B, 3, 31, 27)
    for b in range(B):
        y0, x0 = int(torch.randint(0, 10, (1,), generator=g)), int(torch.randint(0, 8, (1,), generator=g))
        imgs[b, :, y0:y0 + 12 + seed, x0:x0 + 9 + 2 * seed] = torch.rand(3, 12 + seed, 9 + 2 * seed, generator=g)
    imgs[0, 0, 30, 26] = 0.0301 if seed == 1 else 0.0                 # a pixel just above the 0.03 threshold in the far corner
    feats = [torch.randn(B, 4, 8, 6, generator=g).requires_grad_(), torch.randn(B, 9, 3, 3, generator=g).requires_grad_()]
    pts = (torch.randn(B, N, 3, generator=g) * 0.3).requires_grad_()
    rb = R.GCNModel.get_bound_of_images(imgs)
    eq(O.image_bounds(imgs), rb, "bounds")
    ref = R.GCNModel.perceptual_feature_pooling(feats, pts, rb)
    got = O.perceptual_feature_pooling(feats, pts, rb)
    close(got, ref, what="pooled features")
    up = torch.randn(ref.shape, generator=g)
    gr = torch.autograd.grad((ref * up).sum(), [pts] + feats)
    go = torch.autograd.grad((got * up).sum(), [pts] + feats)
    close(go[0], gr[0], rtol=1e-4, atol=1e-4, what="grad points")
    close(go[1], gr[1], what="grad feat 0"); close(go[2], gr[2], what="grad feat 1")


def test_meshing(R):
    g = torch.Generator().manual_seed(7)
    B = 2
    v = (torch.sigmoid(torch.randn(B, 3, generator=g)) + 0.1) / torch.tensor([8.0, 10.0, 10.0])
    q, t = torch.sigmoid(torch.randn(B, 4, generator=g)), torch.tanh(torch.randn(B, 3, generator=g))
    from vpn_b200 import templates
    for kind, fn in (("sphere", R.Meshing.sphere_meshing), ("cuboid", R.Meshing.cuboid_meshing)):
        ref = torch.stack([m.vertices for m in fn(v, q, t)])
        tv, _ = templates.template(kind, "cpu")
        eq(O.mesh_vertices(tv, v, q, t), ref, kind + " mesh vertices")
