"""f4: the two custom ops are differentiable like the reference's (op/upfirdn2d.py:20-146, op/fused_act.py:20-84): first and
second derivatives on the GPU against autograd through the CPU oracle's pure-torch restatements."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("up,down,pad,ksize", [(1, 1, (1, 1), 4), (2, 1, (2, 1), 4), (1, 2, (1, 1), 4), ((2, 1), (1, 2), (0, 1, 2, 0), 3)])
def test_upfirdn2d_grad_and_gradgrad(up, down, pad, ksize):
    from oracle import vt_oracle as O
    from vtoonify_b200.op import upfirdn2d
    with torch.enable_grad():
        x_c = _rand((2, 3, 9, 11), 1).requires_grad_(True)
        k = torch.rand((ksize, ksize), generator=torch.Generator().manual_seed(2)) + 0.1
        k = k / k.sum()
        w = None
        y_c = O.upfirdn2d(x_c, k, up=up, down=down, pad=pad)
        w = _rand(tuple(y_c.shape), 3)
        g_c, = torch.autograd.grad((y_c * w).sum(), x_c, create_graph=True)
        v = _rand(tuple(g_c.shape), 4)
        x_g = x_c.detach().cuda().requires_grad_(True)
        y_g = upfirdn2d(x_g, k.cuda(), up=up, down=down, pad=pad)
        assert y_g.requires_grad and tuple(y_g.shape) == tuple(y_c.shape)
        assert (y_g.detach().cpu() - y_c.detach()).abs().max().item() <= 1e-5
        w_g = w.cuda().requires_grad_(True)
        g_g, = torch.autograd.grad((y_g * w_g).sum(), x_g, create_graph=True)
        assert (g_g.detach().cpu() - g_c.detach()).abs().max().item() <= 1e-5
        # second derivative: d/dw of <grad_x, v> = upfirdn2d(v)  (UpFirDn2dBackward.backward)
        ggw_g, = torch.autograd.grad((g_g * v.cuda()).sum(), w_g)
        ref = O.upfirdn2d(v, k, up=up, down=down, pad=pad)
        assert (ggw_g.cpu() - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("shape,has_bias", [((2, 8, 5, 7), True), ((3, 16), True), ((2, 4, 6, 6), False)])
def test_fused_leaky_relu_grad_and_gradgrad(shape, has_bias):
    from oracle import vt_oracle as O
    from vtoonify_b200.op import FusedLeakyReLU, fused_leaky_relu
    with torch.enable_grad():
        x_c = _rand(shape, 5).requires_grad_(True)
        b_c = (_rand((shape[1],), 6) * 0.3).requires_grad_(True) if has_bias else None
        y_c = O.fused_leaky_relu(x_c, b_c, 0.2, 2 ** 0.5)
        w = _rand(shape, 7)
        ins_c = [x_c] + ([b_c] if has_bias else [])
        grads_c = torch.autograd.grad((y_c * w).sum(), ins_c, create_graph=True)
        x_g = x_c.detach().cuda().requires_grad_(True)
        b_g = b_c.detach().cuda().requires_grad_(True) if has_bias else None
        y_g = fused_leaky_relu(x_g, b_g, 0.2, 2 ** 0.5)
        assert torch.equal(y_g.detach().cpu(), y_c.detach())
        w_g = w.cuda().requires_grad_(True)
        ins_g = [x_g] + ([b_g] if has_bias else [])
        grads_g = torch.autograd.grad((y_g * w_g).sum(), ins_g, create_graph=True)
        for a, b in zip(grads_g, grads_c):
            assert (a.detach().cpu() - b.detach()).abs().max().item() <= 1e-5 * max(1.0, b.detach().abs().max().item())
        # double backward (R1-style penalties differentiate the gradient): d/dw of <grad_x, v> (+ <grad_b, u>)
        v = _rand(shape, 8)
        u = _rand((shape[1],), 9)
        obj_c = (grads_c[0] * v).sum() + ((grads_c[1] * u).sum() if has_bias else 0.0)
        obj_g = (grads_g[0] * v.cuda()).sum() + ((grads_g[1] * u.cuda()).sum() if has_bias else 0.0)
        w_leaf = w.clone().requires_grad_(True)
        y_c2 = O.fused_leaky_relu(x_c, b_c, 0.2, 2 ** 0.5)
        g2 = torch.autograd.grad((y_c2 * w_leaf).sum(), ins_c, create_graph=True)
        obj_c2 = (g2[0] * v).sum() + ((g2[1] * u).sum() if has_bias else 0.0)
        ggw_c, = torch.autograd.grad(obj_c2, w_leaf)
        ggw_g, = torch.autograd.grad(obj_g, w_g)
        assert (ggw_g.cpu() - ggw_c).abs().max().item() <= 1e-5 * max(1.0, ggw_c.abs().max().item())
        del obj_c
        # the module form
        m = FusedLeakyReLU(shape[1], bias=has_bias).cuda()
        if has_bias:
            m.bias.data.copy_(b_c.detach().cuda())
        ym = m(x_g)
        assert torch.equal(ym.detach().cpu(), y_c.detach()) and ym.requires_grad
    with torch.no_grad():
        assert not fused_leaky_relu(x_g.detach(), None if b_g is None else b_g.detach()).requires_grad
