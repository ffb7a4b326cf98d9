"""CPU: the BatchNorm folding behind the evaluation GEMMs (ops.FoldedMLP: W' = W * gamma / sqrt(var + eps),
b' = (b - mean) * gamma / sqrt(var + eps) + beta, zero padded to multiples of 8) against conv -> BatchNorm(eval) in fp32."""
import torch
import torch.nn as nn

from pointcloud_bridge_b200 import ops


def test_folded_weights_reproduce_conv_bn_eval():
    torch.manual_seed(0)
    widths = [13, 20, 196, 5]
    convs = [nn.Conv2d(a, b, 1, bias=(i != 1)) for i, (a, b) in enumerate(zip(widths[:-1], widths[1:]))]
    bns = [nn.BatchNorm2d(b) for b in widths[1:-1]] + [None]                   # last layer: plain conv (a classifier)
    for bn in bns[:-1]:
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5), bn.bias.normal_(0, 0.3), bn.running_mean.normal_(0, 0.3), bn.running_var.uniform_(0.5, 2)
        bn.eval()
    f = ops.FoldedMLP(convs, bns)
    assert [tuple(w.shape) for w in f.w] == [(24, 16), (200, 24), (8, 200)] and f.n == [20, 196, 5] and f.cout == 5
    x = torch.randn(64, 13)
    ref = x
    for conv, bn in zip(convs, bns):
        ref = torch.nn.functional.linear(ref, conv.weight.flatten(1), conv.bias)
        if bn is not None:
            ref = torch.relu(torch.nn.functional.batch_norm(ref, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps))
    h = torch.nn.functional.pad(x, (0, 3))
    for i, (w, b) in enumerate(zip(f.w, f.b)):
        assert w.dtype == torch.bfloat16 and b.dtype == torch.float32
        # exact folded weights (before the bf16 rounding the kernel's operands get): pad rows / columns are zero
        wf = torch.zeros_like(w, dtype=torch.float32)
        conv, bn = convs[i], bns[i]
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps) if bn is not None else torch.ones(conv.weight.shape[0])
        wf[:conv.weight.shape[0], :conv.weight.shape[1]] = conv.weight.flatten(1) * scale[:, None]
        assert torch.equal(w, wf.to(torch.bfloat16))
        assert float(b[f.n[i]:].abs().sum()) == 0.0
        h = h[:, :w.shape[1]] @ wf.t() + b
        if i < 2:
            h = torch.relu(h)
    assert torch.allclose(h[:, :5], ref, rtol=1e-4, atol=1e-5)
    assert float(h[:, 5:].abs().max()) == 0.0
