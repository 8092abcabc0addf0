"""The oracle is pinned against vectors produced by the REFERENCE's own code (oracle/make_golden.py)."""
import numpy as np
import torch

from oracle import vpn_oracle as O

T = lambda a: torch.from_numpy(np.asarray(a))


def eq(a, b):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_array_equal(a, b)


def close(a, b, rtol=1e-6, atol=1e-7):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def test_transform_bit_exact(golden):
    g = golden
    pts, q, t = T(g["in_tf_points"]), T(g["in_tf_q"]), T(g["in_tf_t"])
    eq(O.refine_quaternions(q), g["ref_tf_refine"])
    eq(O.rotation_matrices(O.refine_quaternions(q)), g["ref_tf_matrices"])
    eq(O.rotate_points(pts, q), g["ref_tf_rotate"])
    eq(O.transform_points(pts, q, t), g["ref_tf_transform"])
    d, e, a, ang = (T(g["in_tf_" + k]) for k in ("dists", "elevs", "azims", "angles"))
    eq(O.view_to_obj_points(pts, d, e, a, ang), g["ref_tf_view_to_obj"])
    eq(O.obj_to_view_points(pts, d, e, a), g["ref_tf_obj_to_view"])
    eq(O.rotate_points_forward_x_axis(pts, ang), g["ref_tf_rotate_x"])


def test_transform_gradients(golden):
    g = golden
    pts, q, t = (T(g["in_tf_" + k]).clone().requires_grad_() for k in ("points", "q", "t"))
    (O.transform_points(pts, q, t) * T(g["in_tf_upstream"])).sum().backward()
    close(pts.grad, g["ref_tf_grad_points"]); close(q.grad, g["ref_tf_grad_q"], 1e-5, 1e-6); close(t.grad, g["ref_tf_grad_t"])


def test_sphere_sampling(golden):
    g = golden
    v, q, t, ue, ua = (T(g["in_sp_" + k]) for k in ("v", "q", "t", "ue", "ua"))
    eq(O.sphere_canonical(v, ue, ua), g["ref_sp_canonical"])
    eq(O.sphere_sampling(v, q, t, ue, ua), g["ref_sp_points"])
    v, q, t = (x.clone().requires_grad_() for x in (v, q, t))
    (O.sphere_sampling(v, q, t, ue, ua) * T(g["in_sp_upstream"])).sum().backward()
    close(v.grad, g["ref_sp_grad_v"], 1e-5, 1e-6); close(q.grad, g["ref_sp_grad_q"], 1e-5, 1e-6); close(t.grad, g["ref_sp_grad_t"], 1e-5, 1e-6)


def test_cuboid_sampling(golden):
    g = golden
    v, q, t, u = (T(g["in_cb_" + k]) for k in ("v", "q", "t", "u"))
    eq(O.cuboid_face_counts(v, u.shape[1]), g["ref_cb_counts"])
    eq(O.cuboid_canonical(v, u), g["ref_cb_canonical"])
    eq(O.cuboid_sampling(v, q, t, u), g["ref_cb_points"])
    eq(O.cuboid_face_counts(T(g["in_cb_counts_v"]), 1000), g["ref_cb_counts_1000"])
    eq(O.cuboid_face_counts(T(g["in_cb_counts_v"]), 4096), g["ref_cb_counts_4096"])
    v, q, t = (x.clone().requires_grad_() for x in (v, q, t))
    (O.cuboid_sampling(v, q, t, u) * T(g["in_cb_upstream"])).sum().backward()
    close(v.grad, g["ref_cb_grad_v"], 1e-5, 1e-6); close(q.grad, g["ref_cb_grad_q"], 1e-5, 1e-6); close(t.grad, g["ref_cb_grad_t"], 1e-5, 1e-6)


def test_chamfer_dense_and_nn(golden, c_oracle):
    g = golden
    p1, p2 = T(g["in_cd_p1"]), T(g["in_cd_p2"])
    eq(O.chamfer_dense(p1, p2), g["ref_cd_loss"])
    eq(O.chamfer_dense(p1, p2, each_batch=True), g["ref_cd_loss_each"])
    eq(O.chamfer_dense(p1, p2, w1=0.5, w2=1.0), g["ref_cd_loss_w"])
    # arg-mins: bit exact.  min VALUES: the golden ones come from torch-CPU sqrt (MKL VML, 1 ulp off on
    # ~0.7% of inputs); the oracle pins IEEE sqrt, which is what the reference computes on its own
    # device ('cuda', config.py:2) - so values are compared to 1 ulp and the C and torch oracles to 0.
    for row_block in (7, 64, 4096):
        m1, i1, m2, i2 = O.chamfer_nn(p1, p2, row_block=row_block)
        eq(i1, g["ref_cd_idx1"]); eq(i2, g["ref_cd_idx2"])
        close(m1, g["ref_cd_min1"], 1.3e-7, 0); close(m2, g["ref_cd_min2"], 1.3e-7, 0)
    c1, ci1, c2, ci2 = c_oracle(g["in_cd_p1"], g["in_cd_p2"])
    eq(m1, c1); eq(i1, ci1); eq(m2, c2); eq(i2, ci2)
    c1, ci1, c2, ci2 = c_oracle(g["in_cd_p1"], g["in_cd_p2"], threads=3)
    eq(m1, c1); eq(i1, ci1); eq(m2, c2); eq(i2, ci2)


def test_chamfer_gradients(golden):
    g = golden
    p1, p2 = T(g["in_cd_p1"]).clone().requires_grad_(), T(g["in_cd_p2"]).clone().requires_grad_()
    O.chamfer_dense(p1, p2).backward()
    close(p1.grad, g["ref_cd_grad_p1"], 1e-5, 1e-8); close(p2.grad, g["ref_cd_grad_p2"], 1e-5, 1e-8)
    # closed form from the arg-mins == autograd through the dense graph
    with torch.no_grad():
        m1, i1, m2, i2 = O.chamfer_nn(p1, p2)
        b, p, m = p1.shape[0], p1.shape[1], p2.shape[1]
        g1 = torch.full((b, p), 1.0 / (p * b)); g2 = torch.full((b, m), 1.0 / (m * b))
        gp1, gp2 = O.chamfer_grad_from_nn(p1, p2, m1, i1, m2, i2, g1, g2)
    close(gp1, g["ref_cd_grad_p1"], 1e-4, 1e-8); close(gp2, g["ref_cd_grad_p2"], 1e-4, 1e-8)


def test_chamfer_round2_golden(golden2, c_oracle):
    """P = 1536, M = 640 clouds (clusters vs box surface; lattice + duplicates straddling chunk borders): the oracle's
    arg-mins equal the reference's torch.min bit for bit, loss and gradients equal the reference's autograd."""
    g = golden2
    p1, p2 = T(g["in_cd_p1"]), T(g["in_cd_p2"])
    eq(O.chamfer_dense(p1, p2), g["ref_cd_loss"])
    eq(O.chamfer_dense(p1, p2, each_batch=True), g["ref_cd_loss_each"])
    m1, i1, m2, i2 = O.chamfer_nn(p1, p2)
    eq(i1, g["ref_cd_idx1"]); eq(i2, g["ref_cd_idx2"])
    close(m1, g["ref_cd_min1"], 1.3e-7, 0); close(m2, g["ref_cd_min2"], 1.3e-7, 0)
    c1, ci1, c2, ci2 = c_oracle(g["in_cd_p1"], g["in_cd_p2"])
    eq(m1, c1); eq(i1, ci1); eq(m2, c2); eq(i2, ci2)
    a, b = p1.clone().requires_grad_(), p2.clone().requires_grad_()
    O.chamfer_dense(a, b).backward()
    close(a.grad, g["ref_cd_grad_p1"], 1e-5, 1e-9); close(b.grad, g["ref_cd_grad_p2"], 1e-5, 1e-9)


def test_view_transform_backward_round2(golden, golden2):
    g = golden
    pts = T(g["in_tf_points"])
    d, e, a, ang = (T(g["in_tf_" + k]) for k in ("dists", "elevs", "azims", "angles"))
    w = T(g["in_tf_upstream"])
    pg = pts.clone().requires_grad_()
    (O.obj_to_view_points(pg, d, e, a) * w).sum().backward()
    close(pg.grad, golden2["ref_tf_obj_to_view_grad_points"], 1e-5, 1e-6)
    pg = pts.clone().requires_grad_()
    (O.rotate_points_forward_x_axis(pg, ang) * w).sum().backward()
    close(pg.grad, golden2["ref_tf_rotate_x_grad_points"], 1e-5, 1e-6)


def test_vp_diverse(golden):
    g = golden
    tr = T(g["in_vd_translates"])
    eq(O.vp_diverse([tr[:, i] for i in range(tr.shape[1])], T(g["in_vd_gt"])), g["ref_vd_loss"])


def test_end_to_end(golden):
    for kind in ("sphere", "cuboid"):
        g = golden
        v, q, t = (T(g[f"in_e2e_{kind}_{k}"]).clone().requires_grad_() for k in ("v", "q", "t"))
        pts = O.sample_predict_points(kind, v, q, t, T(g[f"in_e2e_{kind}_u"]))
        eq(pts, g[f"ref_e2e_{kind}_points"])
        loss = O.chamfer_dense(pts, T(g[f"in_e2e_{kind}_target"]))
        eq(loss, g[f"ref_e2e_{kind}_loss"])
        loss.backward()
        for k, x in (("v", v), ("q", q), ("t", t)):
            close(x.grad, g[f"ref_e2e_{kind}_grad_{k}"], 1e-5, 1e-8)


def test_meshing(golden, golden_templates):
    g, tm = golden, golden_templates
    v, q, t = T(g["in_ms_v"]), T(g["in_ms_q"]), T(g["in_ms_t"])
    sph = O.sphere_template(T(tm["sphere_vertices"]))
    cub = T(tm["cuboid_vertices"])
    sv, cv = O.mesh_vertices(sph, v, q, t), O.mesh_vertices(cub, v, q, t)
    eq(sv, g["ref_ms_sphere_vertices"]); eq(cv, g["ref_ms_cuboid_vertices"])
    sf, cf = T(tm["sphere_faces"]).long(), T(tm["cuboid_faces"]).long()
    cvs, cfs = O.compose_meshes([sv[0], cv[0], sv[1]], [sf, cf, sf])
    eq(cvs, g["ref_ms_compose_vertices"]); eq(cfs, g["ref_ms_compose_faces"])
    # facts recorded in SURVEY.md section 8c
    assert tm["sphere_vertices"].shape == (128, 3) and tm["sphere_faces"].shape == (252, 3)
    assert tm["cuboid_vertices"].shape == (128, 3) and tm["cuboid_faces"].shape == (504, 3)
    assert tm["sphere386_vertices"].shape == (386, 3) and tm["sphere386_faces"].shape == (768, 3)


def test_silhouette_oracle_sanity(golden_templates):
    """Render parity is UNPINNED (kaolin absent); this only checks the restatement's invariants."""
    tm = golden_templates
    sph = O.sphere_template(T(tm["sphere_vertices"]))
    faces = T(tm["sphere_faces"]).long()
    verts = (sph * 0.25)[None].clone().requires_grad_()
    rot, pos = O.look_at_camera(0.0, 0.0, 1.0)
    alpha = O.soft_silhouette(verts, faces, rot[None], pos[None], 32, 32)
    a = alpha.detach()
    assert a.min() >= 0 and a.max() == 1.0
    assert a[0, 16, 16] == 1.0 and a[0, 0, 0] < 1e-6           # centre covered, corner empty
    assert ((a > 0) & (a < 1)).sum() > 0                         # soft rim exists
    alpha.sum().backward()
    assert torch.isfinite(verts.grad).all() and verts.grad.abs().sum() > 0
    # camera sits at (1,0,0) for dist=1, elev=azim=0
    np.testing.assert_allclose(pos.numpy(), [1, 0, 0], atol=1e-7)


def test_silhouette_oracle_cull_modes_and_pixel_subset(golden_templates):
    """soft_cull_backfaces: on a closed outward-wound mesh the back faces lie behind covered pixels or on the rim, so
    culling them in the soft pass can only lower alpha; a mesh seen from inside out (all faces back-facing) is
    invisible when culled and leaves a soft-only image otherwise.  `pixels=` returns exactly the full image's entries."""
    tm = golden_templates
    sph = O.sphere_template(T(tm["sphere_vertices"]))
    faces = T(tm["sphere_faces"]).long()
    verts = (sph * 0.3)[None] + torch.tensor([0.0, 0.1, -0.05])
    rot, pos = O.look_at_camera(30.0, 20.0, 1.4)
    full = O.soft_silhouette(verts, faces, rot[None], pos[None], 40, 40)
    cull = O.soft_silhouette(verts, faces, rot[None], pos[None], 40, 40, soft_cull_backfaces=True)
    assert ((full == 1.0) == (cull == 1.0)).all()                     # coverage pass is the same
    assert (full >= cull - 1e-7).all() and (full > cull + 1e-6).any()
    inward = faces[:, [0, 2, 1]]
    inv_cull = O.soft_silhouette(verts, inward, rot[None], pos[None], 40, 40, soft_cull_backfaces=True)
    inv_full = O.soft_silhouette(verts, inward, rot[None], pos[None], 40, 40)
    # inside-out sphere: the far hemisphere now faces the camera (covers the disc); culled soft pass still finds it
    assert inv_cull.max() == 1.0 and inv_full.max() == 1.0
    pix = torch.tensor([0, 41, 20 * 40 + 20, 13 * 40 + 9, 39 * 40 + 39, 17 * 40 + 30])
    for mode in (False, True):
        sub = O.soft_silhouette(verts, faces, rot[None], pos[None], 40, 40, soft_cull_backfaces=mode, pixels=pix)
        ref = O.soft_silhouette(verts, faces, rot[None], pos[None], 40, 40, soft_cull_backfaces=mode).reshape(1, -1)[:, pix]
        eq(sub, ref.numpy())


def test_mesh_sample_oracle_properties(golden_templates):
    """kaolin's TriangleMesh.sample is absent (parity unpinned): pin the restatement's own invariants."""
    from oracle import vpn_oracle as O
    verts = torch.from_numpy(golden_templates["sphere386_vertices"]).float()
    faces = torch.from_numpy(golden_templates["sphere386_faces"]).long()
    cdf = O.mesh_face_cdf(verts.numpy(), faces.numpy())
    assert cdf.dtype == np.float32 and np.all(np.diff(cdf) >= 0) and abs(float(cdf[-1]) - 1.0) < 1e-6
    # triangle areas against a float64 cross product
    v = verts.double().numpy(); f = faces.numpy()
    area = 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1)
    np.testing.assert_allclose(np.diff(np.concatenate([[0.0], cdf.astype(np.float64)])), area / area.sum(), rtol=2e-4, atol=1e-7)
    g = torch.Generator().manual_seed(0)
    u = torch.rand(50000, 3, generator=g)
    vt = verts.clone().requires_grad_()
    pts, face = O.mesh_sample(vt, faces, u)
    assert pts.shape == (50000, 3) and face.dtype == torch.int64 and int(face.min()) >= 0 and int(face.max()) < faces.shape[0]
    # u0 = 0 -> face 0; sqrt(u1) = 0 -> vertex 0 of the face
    p0, f0 = O.mesh_sample(verts, faces, torch.tensor([[0.0, 0.0, 0.3]]))
    assert int(f0[0]) == 0 and torch.equal(p0[0], verts[faces[0, 0]])
    pts.sum().backward()
    np.testing.assert_allclose(float(vt.grad.sum()), 3 * 50000.0, rtol=1e-5)       # barycentric weights sum to 1
    freq = np.bincount(face.numpy(), minlength=faces.shape[0]) / 50000.0
    share = area / area.sum()
    assert np.abs(freq - share).max() < 5 * np.sqrt(share.max() / 50000.0)


def test_emd_auction_oracle_properties():
    """The reference's auction is non-deterministic (racing writes) and cannot be run here: the restatement is
    pinned by its own invariants - a valid, near-optimal assignment - against scipy's exact solver."""
    from scipy.optimize import linear_sum_assignment
    from oracle import vpn_oracle as O
    rng = np.random.default_rng(5)
    for n in (1, 7, 64, 300):
        a = rng.random((n, 3), dtype=np.float32); b = rng.random((n, 3), dtype=np.float32)
        dist, ass = O.emd_auction(a, b, 0.005, 50)
        assert ass.dtype == np.int32 and ass.min() >= 0 and ass.max() < n
        np.testing.assert_allclose(dist, ((a - b[ass]) ** 2).sum(1), rtol=1e-6, atol=1e-9)
        cost = np.linalg.norm(a[:, None] - b[None], axis=2)
        r, c = linear_sum_assignment(cost)
        opt = cost[r, c].mean()
        got = np.sqrt(dist).mean()
        # an eps-auction is within n*eps of optimal once everything is assigned; the forced last round can only lower it
        assert got <= opt + 0.005 * 1.5 + 1e-6, (n, got, opt)
        assert len(set(ass.tolist())) >= 0.9 * n                       # almost a bijection
        d2, a2 = O.emd_auction(a, b, 0.005, 50)
        assert (a2 == ass).all() and (d2 == dist).all()                 # deterministic
    # identical clouds: every bidder prefers its own twin; the assignment is the identity
    a = rng.random((128, 3), dtype=np.float32)
    dist, ass = O.emd_auction(a, a.copy(), 0.005, 50)
    assert (ass == np.arange(128)).all() and (dist == 0).all()
    g = O.emd_backward(a, a[::-1].copy(), ass, np.ones(128, np.float32))
    np.testing.assert_allclose(g, 2.0 * (a - a[::-1][ass]), rtol=1e-6)


def test_gcn_feature_pooling(golden):
    """gcn.py:84-164 (image bounds + perceptual feature pooling): the restatement against the reference's own output."""
    g = golden
    imgs, pts = T(g["in_pool_imgs"]), T(g["in_pool_points"])
    feats = [T(g[f"in_pool_feat{i}"]) for i in range(3)]
    bounds = O.image_bounds(imgs)
    eq(bounds, g["ref_pool_bounds"])
    fg = [f.clone().requires_grad_() for f in feats]
    pg = pts.clone().requires_grad_()
    out = O.perceptual_feature_pooling(fg, pg, bounds)
    close(out, g["ref_pool_out"], 1e-5, 1e-6)
    (out * T(g["in_pool_upstream"])).sum().backward()
    close(pg.grad, g["ref_pool_grad_points"], 1e-4, 1e-4)
    for i, f in enumerate(fg):
        close(f.grad, g[f"ref_pool_grad_feat{i}"], 1e-5, 1e-6)
    close(O.perceptual_feature_pooling(feats, pts, O.image_bounds(imgs)), g["ref_pool_local"], 1e-5, 1e-6)
    # the `== 0` quirk of gcn.py:107: an occupied first column is skipped as a lower bound
    im = torch.zeros(1, 3, 8, 8); im[0, :, 2:5, 0] = 1.0; im[0, :, 2:5, 3] = 1.0
    b = O.image_bounds(im)
    assert float(b[0, 0]) == 3 / 8 * 2 - 1 and float(b[0, 1]) == 3 / 8 * 2 - 1
    eq(O.image_bounds(torch.zeros(1, 3, 8, 6)), np.array([[-1.0, 1.0, -1.0, 1.0]], np.float32))
