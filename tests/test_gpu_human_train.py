"""GPU parity of the human trainer's forward / adjoint kernels through the C ABI
(neuman_b200.autograd / ops -> libneuman_b200.so) against torch autograd of the oracle restatement on the CPU
(oracle/neuman_oracle.py: warp_diff_Tinv, eval_human_samples, vertex_forward, validated against the reference's own
functions in tests/test_oracle_vs_reference.py).  The same kernel bodies are checked on the host by
tests/test_human_train_emu.py."""
import copy

import numpy as np
import pytest
import torch

import neuman_b200 as nb
from neuman_b200 import autograd as nag
from neuman_b200 import ops
from neuman_b200 import train as nt
from oracle import neuman_oracle as no
from oracle import synth_smpl

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to(DEV, dtype)


def _rays_through_body(V, seed, R, S):
    rng = np.random.RandomState(seed)
    a = V[rng.randint(0, V.shape[0], R)] + rng.normal(0, 0.05, (R, 3))
    d = rng.normal(0, 1, (R, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    t = np.linspace(-0.12, 0.12, S)
    return (a[:, None] + t[None, :, None] * d[:, None]).astype(np.float32)


def _close(got, want, rel, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err, ref = np.abs(got - want).max(), np.abs(want).max()
    assert err < rel * (1 + ref), (what, err, ref)


def test_warp_diff_and_canonicalize_kernels():
    body = synth_smpl.random_body(seed=3, center=(0.0, 0.1, 0.2))
    V = body["verts"].astype(np.float32)
    F = np.ascontiguousarray(np.asarray(body["faces"])[:, :3])
    T = body["Ts"][:V.shape[0]].astype(np.float32)
    R, S = 96, 32
    P = _rays_through_body(V, 3, R, S)
    rng = np.random.RandomState(5)
    off = rng.normal(0, 0.01, (R, S, 3)).astype(np.float32)
    verts, Tt, offt = cu(V).requires_grad_(True), cu(T).requires_grad_(True), cu(off).requires_grad_(True)
    # the query once (constants of the step), shared by the device path and the oracle
    sd, f_id, closest = nb.signed_distance(cu(P.reshape(-1, 3)), verts.detach(), F)
    I, Cl = f_id.cpu().numpy(), closest.cpu().numpy()
    Vo, To = torch.from_numpy(V).requires_grad_(True), torch.from_numpy(T).requires_grad_(True)
    offo = torch.from_numpy(off).requires_grad_(True)
    # ---- drop-in form: T_interp_inv and its adjoint ----
    Ti = nag.warp_diff_tinv(verts, Tt, f_id, closest, F)
    Ti_o = no.warp_diff_Tinv(Cl, I, Vo, F, To)
    _close(Ti.detach().cpu(), Ti_o.detach(), 2e-6, "Tinv")
    w = rng.normal(0, 1, (R * S, 4, 4)).astype(np.float32)
    (Ti * cu(w)).sum().backward()
    (Ti_o * torch.from_numpy(w)).sum().backward()
    _close(Tt.grad.cpu(), To.grad, 2e-4, "dT")
    _close(verts.grad.cpu(), Vo.grad, 2e-4, "dverts")
    # ---- fused form ----
    verts.grad = Tt.grad = Vo.grad = To.grad = None
    cp, cd = nag.human_canonicalize(cu(P), verts, Tt, f_id, closest, F, offt)
    cp_o, cd_o = no.eval_human_samples(torch.from_numpy(P), Cl, I, Vo, F, To, offo)
    _close(cp.detach().cpu(), cp_o.detach(), 2e-6, "can_pts")
    # unit differences of points 8e-3 apart: the 1-ulp (1e-7) differences of can_pts are amplified by 1 / spacing
    assert np.abs(cd.detach().cpu().numpy() - cd_o.detach().numpy()).max() < 1e-4
    w1, w2 = rng.normal(0, 1, (R, S, 3)).astype(np.float32), rng.normal(0, 1, (R, S, 3)).astype(np.float32)
    ((cp * cu(w1)).sum() + (cd * cu(w2)).sum()).backward()
    ((cp_o * torch.from_numpy(w1)).sum() + (cd_o * torch.from_numpy(w2)).sum()).backward()
    _close(offt.grad.cpu(), offo.grad, 2e-4, "doffset")
    _close(Tt.grad.cpu(), To.grad, 2e-4, "dT fused")
    _close(verts.grad.cpu(), Vo.grad, 2e-4, "dverts fused")
    # only the directions / only the points carry a gradient; no offset
    cp2, cd2 = nag.human_canonicalize(cu(P), verts.detach().requires_grad_(True), Tt.detach(), f_id, closest, F, None)
    assert torch.equal(cd2, nag.human_canonicalize(cu(P), verts.detach(), Tt.detach(), f_id, closest, F, None)[1])
    cd2.sum().backward()
    # ops-level wrapper (query inside) returns the same numbers and the reference's extra outputs
    cp3, cd3, fid3, sd3 = ops.eval_human_samples(cu(P), verts.detach(), F, Tt.detach(), offset=offt.detach())
    assert torch.equal(fid3, f_id) and torch.equal(sd3, sd) and torch.equal(cp3, cp.detach()) and torch.equal(cd3, cd.detach())
    # empty batch
    z = nag.warp_diff_tinv(verts.detach(), Tt.detach(), f_id[:0], closest[:0], F)
    assert z.shape == (0, 4, 4)


def _model_and_params(seed=4):
    model = synth_smpl.torch_model(0)
    nj, nb_ = model["parents"].shape[0], model["shapedirs"].shape[-1]
    rng = np.random.RandomState(seed)
    pose = rng.normal(0, 0.3, (1, 3 * nj)).astype(np.float32)
    betas = rng.normal(0, 1.0, (1, nb_)).astype(np.float32)
    ang = 0.2
    align = np.eye(4, dtype=np.float32)
    align[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    align = align.T.copy()
    align[3, :3] = (0.3, -0.1, 2.0)
    return model, pose, betas, align


def _device_model(model):
    return ops.SmplModelDevice(model["v_template"], model["shapedirs"], model["J_regressor"], model["weights"],
                               model["parents"], device=DEV)


def test_vertex_forward_training_kernels_and_adjoint():
    model, pose, betas, align = _model_and_params()
    scale = 0.4
    dm = _device_model(model)
    nv = dm.n_verts
    p, b, a = (cu(x).requires_grad_(True) for x in (pose, betas, align))
    da = no.da_pose(dm.n_joints)
    world, T = nag.vertex_forward(dm, p, b, a, scale, da.to(DEV))
    po, bo, ao = (torch.from_numpy(x).requires_grad_(True) for x in (pose, betas, align))
    world_o, T_o = no.vertex_forward(model, po, bo, ao, scale)
    assert world.shape == (1, nv, 3) and T.shape == (1, nv, 4, 4)
    _close(T.detach().cpu(), T_o.detach(), 5e-6, "T_da2scene")
    _close(world.detach().cpu(), world_o.detach(), 5e-6, "world_verts")
    rng = np.random.RandomState(8)
    gT, gw = rng.normal(0, 1, (1, nv, 4, 4)).astype(np.float32), rng.normal(0, 1, (1, nv, 3)).astype(np.float32)
    ((T * cu(gT)).sum() + (world * cu(gw)).sum()).backward()
    ((T_o * torch.from_numpy(gT)).sum() + (world_o * torch.from_numpy(gw)).sum()).backward()
    _close(p.grad.cpu(), po.grad, 5e-4, "dpose")
    _close(b.grad.cpu(), bo.grad, 5e-4, "dbetas")
    _close(a.grad.cpu(), ao.grad, 5e-4, "dalignment")
    # each output alone
    for which in (0, 1):
        p.grad = b.grad = a.grad = po.grad = bo.grad = ao.grad = None
        out, out_o = nag.vertex_forward(dm, p, b, a, scale, da.to(DEV)), no.vertex_forward(model, po, bo, ao, scale)
        (out[which] * cu((gw, gT)[which])).sum().backward()
        (out_o[which] * torch.from_numpy((gw, gT)[which])).sum().backward()
        _close(p.grad.cpu(), po.grad, 5e-4, f"dpose[{which}]")
        _close(a.grad.cpu(), ao.grad, 5e-4, f"dalignment[{which}]")
    # the inference path (float64 algebra, nm_smpl_scene_transforms) agrees with the training forward
    w_inf, _, T_inf = ops.smpl_scene_transforms(dm, cu(pose), cu(betas), align, scale)
    _close(T.detach().cpu()[0], T_inf[:nv].float().cpu(), 5e-6, "training vs inference T")
    _close(world.detach().cpu()[0], w_inf.cpu(), 5e-6, "training vs inference world")


def test_eval_human_samples_end_to_end():
    """neuman_b200.train.eval_human_samples (trainers/human_nerf_trainer.py:241-278): rays -> samples -> offset net ->
    vertex_forward -> closest-face query -> canonical points / directions -> human net, and loss.backward() down to
    poses / betas / alignments / offset-net weights, against the oracle chain on the CPU."""
    from neuman_b200.synthetic import make_model
    model, pose, betas, align = _model_and_params(seed=6)
    scale = 0.4
    torch.manual_seed(2)
    opt = nb.default_opt(use_cuda=True, num_offset_nets=1, offset_scale=0.02, offset_scale_type='tanh', samples_per_ray=24)
    net = nb.HumanNeRF(opt, poses=pose, betas=betas, alignments=align[None], scale=scale, smpl_model=make_model(0))
    F = np.ascontiguousarray(model["faces"][:, :3])
    with torch.no_grad():
        world0, _ = net.vertex_forward(0)
    V0 = world0[0].cpu().numpy()
    # rays from a pinhole in front of the body through its vertices
    rng = np.random.RandomState(3)
    R = 64
    eye = V0.mean(0) + np.array([0.0, 0.0, -2.0])
    tgt = V0[rng.randint(0, V0.shape[0], R)] + rng.normal(0, 0.01, (R, 3))
    d = tgt - eye
    dist = np.linalg.norm(d, axis=1, keepdims=True)
    d = (d / dist).astype(np.float32)
    batch = {'origin': cu(np.repeat(eye[None], R, 0)), 'direction': cu(d), 'human_near': cu(dist - 0.15),
             'human_far': cu(dist + 0.15), 'cur_view_f': 3 / 11, 'cap_id': 0}
    human_pts, human_dirs, z, can_pts, can_dirs, out = nt.eval_human_samples(net, batch, opt, F, offset_net=net.offset_nets[0])
    S = opt.samples_per_ray
    assert human_pts.shape == (R * S, 3) and can_pts.shape == (R, S, 3) and out.shape == (R, S, 4)
    # ---- oracle chain on the CPU ----
    po, bo = torch.from_numpy(pose).requires_grad_(True), torch.from_numpy(betas).requires_grad_(True)
    ao = torch.from_numpy(align).requires_grad_(True)
    world_o, T_o = no.vertex_forward(model, po, bo, ao, scale)
    off_cpu = copy.deepcopy(net.offset_nets[0]).cpu()
    pts_cpu = human_pts.detach().cpu().reshape(R, S, 3)
    offset_o = off_cpu(torch.cat([pts_cpu, torch.ones(R, S, 1) * (3 / 11)], -1))
    with torch.enable_grad():
        mesh_t, _ = net.vertex_forward(0)                     # the float32 training forward the step itself used
    sd, f_id, closest = nb.signed_distance(human_pts.detach(), mesh_t[0].detach(), F)
    cp_o, cd_o = no.eval_human_samples(pts_cpu, closest.cpu().numpy(), f_id.cpu().numpy(), world_o[0], F, T_o[0], offset_o)
    _close(can_pts.detach().cpu(), cp_o.detach(), 5e-6, "can_pts")
    assert np.abs(can_dirs.detach().cpu().numpy() - cd_o.detach().numpy()).max() < 1e-4
    hp = no.net_params_from_joiner(copy.deepcopy(net.coarse_human_net).cpu())
    out_o = no.net_forward(hp, cp_o, cd_o)
    assert np.abs(out.detach().cpu().numpy() - out_o.detach().numpy()).max() < 2e-3        # tensor-core operands (11 bits)
    # ---- gradients: tight through the geometric chain, loose through the fp16-operand network ----
    w1, w2 = rng.normal(0, 1, (R, S, 3)).astype(np.float32), rng.normal(0, 1, (R, S, 3)).astype(np.float32)
    ((can_pts * cu(w1)).sum() + (can_dirs * cu(w2)).sum()).backward()
    ((cp_o * torch.from_numpy(w1)).sum() + (cd_o * torch.from_numpy(w2)).sum()).backward(retain_graph=True)
    _close(net.poses.grad.cpu()[0], po.grad[0], 1e-3, "dposes")
    _close(net.betas.grad.cpu()[0], bo.grad[0], 1e-3, "dbetas")
    _close(net.alignments.grad.cpu()[0], ao.grad, 1e-3, "dalignments")
    # the offset network runs on the tensor-core kernels (fp16 operands): the Joiner's gradient tolerance (tests/test_gpu_train.py)
    for (k, pm), po_ in zip(net.offset_nets[0].named_parameters(), off_cpu.parameters()):
        g, g_o = pm.grad.cpu().numpy(), po_.grad.numpy()
        assert np.isfinite(g).all() and np.abs(g - g_o).max() < 8e-2 * np.abs(g_o).max() + 1e-7, ("offset net " + k, np.abs(g - g_o).max(), np.abs(g_o).max())
    net.zero_grad()
    po.grad = bo.grad = ao.grad = None
    w3 = rng.normal(0, 1, (R, S, 4)).astype(np.float32)
    # a second loss needs a second forward: the networks' activation stashes are released by the first backward
    out = nt.eval_human_samples(net, batch, opt, F, offset_net=net.offset_nets[0])[5]
    (out * cu(w3)).sum().backward()
    (out_o * torch.from_numpy(w3)).sum().backward()
    g, g_o = net.poses.grad.cpu()[0].numpy(), po.grad[0].numpy()
    assert np.isfinite(g).all() and np.abs(g - g_o).max() < 8e-2 * (1 + np.abs(g_o).max()), (np.abs(g - g_o).max(), np.abs(g_o).max())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.coarse_human_net.parameters())


@pytest.mark.parametrize("scale_type", ["tanh", "linear"])
def test_offset_net_on_the_tensor_core_kernels(scale_type):
    """OffsetNet (models/vanilla.py:169-205) at the step's time as a Joiner on k_mlp_tc / k_mlp_tc_bwd / k_dw_gemm
    (models.offset_forward_at_time) against its own float32 library evaluation: values to the tensor-core tolerance of
    every other network, parameter gradients to the tolerance of the Joiner's (tests/test_gpu_train.py)."""
    from neuman_b200._lib import Context
    opt = nb.default_opt(use_cuda=True, num_offset_nets=1, offset_scale=0.05, offset_scale_type=scale_type)
    torch.manual_seed(7)
    net = nb.build_offset_net(opt)
    n, t = 5000, 3 / 11
    x = torch.randn(n, 3, device=DEV) * 0.8
    x4 = torch.cat([x, torch.full((n, 1), t, device=DEV)], -1)
    lib = net.forward_library(x4)
    l0 = Context.get(0).launch_count()
    tc = net(x4)                                                   # constant time column -> tensor-core path
    assert Context.get(0).launch_count() > l0
    assert (tc - lib).abs().max() < 2e-3 * float(net.nerf.scale) + 1e-6, (tc - lib).abs().max()
    again = net.forward_at_time(x.reshape(50, 100, 3), t)
    assert again.shape == (50, 100, 3) and (again.reshape(-1, 3) - tc).abs().max() < 1e-7
    # a varying time column keeps the library path (bit-equal to forward_library)
    x4v = x4.clone()
    x4v[::2, 3] = 0.5
    assert torch.equal(net(x4v), net.forward_library(x4v))
    # gradients to the offset net's own parameters
    w = torch.randn(n, 3, device=DEV)
    params = list(net.nerf.parameters())
    g_lib = torch.autograd.grad((lib * w).sum(), params)
    g_tc = torch.autograd.grad((net(x4) * w).sum(), params)
    for (k, _), a, b in zip(net.nerf.named_parameters(), g_lib, g_tc):
        assert torch.isfinite(b).all()
        assert (a - b).abs().max() < 8e-2 * (1e-12 + a.abs().max()) + 1e-7, (k, float((a - b).abs().max()), float(a.abs().max()))
    # two evaluations (different times) before one backward: each backward sees the weights of its own forward
    o1, o2 = net.forward_at_time(x, 0.1), net.forward_at_time(x, 0.9)
    g1 = torch.autograd.grad((o1 * w).sum(), params)
    ref1 = torch.autograd.grad((net.forward_library(torch.cat([x, torch.full((n, 1), 0.1, device=DEV)], -1)) * w).sum(), params)
    a, b = ref1[0], g1[0]
    assert (a - b).abs().max() < 8e-2 * a.abs().max() + 1e-7
    # inference
    with torch.no_grad():
        assert (net(x4) - lib).abs().max() < 2e-3 * float(net.nerf.scale) + 1e-6
