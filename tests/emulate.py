"""CPU emulation of the *kernel dataflow* from the packed weights (test helper, not product code).

Follows exactly what the sm_100a kernels compute -- packed operand layouts, the gate-interleaved row
order, the tap-major K order, the folded constants, the deferred skip GEMM and the bf16 rounding points
(h, gate and skip tiles are stored as bf16; accumulation is fp32) -- in plain torch.  Two uses:

* on the CPU (no GPU needed) it checks the packing + reformulation against the oracle;
* on the GPU box it is the sharp comparison for the kernels (same rounding points => tight tolerance).
"""

import torch


def _bf16(x, on):
    return x.to(torch.bfloat16).float() if on else x


def _tf32(x):
    """Nearest tf32, ties away from zero (``audiopure_b200.wavenet.round_to_tf32``)."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def emulate_eps_tf32(packed, x, t, layers, cycle, return_inter=False):
    """The tf32 build: the residual stream, gates and skip sum stay fp32; every GEMM sees its activation operand
    rounded to tf32 (weights are packed already rounded); accumulation is fp32."""
    B, _, L = x.shape
    xs = x[:, 0, :]
    h = torch.relu(xs[..., None] * packed["w0"].float() + packed["b0"].float()) + packed["part0"][t].float()
    gates = []
    inter = {"h": [], "gate": []}
    for n in range(layers):
        d = 2 ** (n % cycle)
        hp = torch.nn.functional.pad(_tf32(h), (0, 0, d, d))
        a = torch.cat([hp[:, tap * d: tap * d + L, :] for tap in range(3)], dim=-1)
        d1 = a @ packed["w1"][n].float().t() + packed["b1"][n].float()
        gate = []
        for c in range(2):
            blk = d1[..., c * 256:(c + 1) * 256]
            gate.append(2.0 * torch.tanh(blk[..., :128]) * torch.sigmoid(2.0 * blk[..., 128:]))  # sigmoid rows are packed halved; the kernels keep 2 x gate (1/2 folded into w2 / ws)
        gate = _tf32(torch.cat(gate, dim=-1))
        gates.append(gate)
        h = h * 0.70710678118654752440 + gate @ packed["w2"][n].float().t() + packed["c2"][t, n].float()
        if return_inter:
            inter["h"].append(h)
            inter["gate"].append(gate)
    g = torch.cat(gates, dim=-1)
    s = _tf32(g @ packed["ws"].float().t() + packed["bs"].float())
    y = torch.relu(s @ packed["wf"].float().t() + packed["bf"].float())
    eps = (y @ packed["wo"].float() + packed["bo"])[:, None, :]
    return (eps, inter) if return_inter else eps


def emulate_eps(packed, x, t, layers, cycle, quantize=True, return_inter=False):
    """x: (B,1,L) fp32 -> eps (B,1,L), from ``WaveNet_Speech_Commands.pack_weights`` output (CPU tensors)."""
    B, _, L = x.shape
    xs = x[:, 0, :]                                                     # (B, L)
    w0, b0 = packed["w0"].float(), packed["b0"].float()
    h = torch.relu(xs[..., None] * w0 + b0) + packed["part0"][t].float()  # (B, L, 256) channels-last
    h = _bf16(h, quantize)
    gates = []
    inter = {"h": [], "gate": []}
    for n in range(layers):
        d = 2 ** (n % cycle)
        w1 = packed["w1"][n].float()                                    # (512, 768), rows permuted, K = tap*256+c
        hp = torch.nn.functional.pad(h, (0, 0, d, d))                   # zero rows outside [0, L)
        a = torch.cat([hp[:, tap * d: tap * d + L, :] for tap in range(3)], dim=-1)  # (B, L, 768)
        d1 = a @ w1.t() + packed["b1"][n].float()                       # (B, L, 512) in packed row order
        gate = []
        for c in range(2):
            blk = d1[..., c * 256:(c + 1) * 256]
            gate.append(2.0 * torch.tanh(blk[..., :128]) * torch.sigmoid(2.0 * blk[..., 128:]))  # sigmoid rows are packed halved; the kernels keep 2 x gate (1/2 folded into w2 / ws)
        gate = _bf16(torch.cat(gate, dim=-1), quantize)                 # (B, L, 256) gate channels in order
        gates.append(gate)
        d2 = gate @ packed["w2"][n].float().t()
        h_next = h * 0.70710678118654752440 + d2 + packed["c2"][t, n].float()
        h = _bf16(h_next, quantize)
        if return_inter:
            inter["h"].append(h)
            inter["gate"].append(gate)
    g = torch.cat(gates, dim=-1)                                        # (B, L, layers*256)
    s = g @ packed["ws"].float().t() + packed["bs"].float()
    s = _bf16(s, quantize)
    y = torch.relu(s @ packed["wf"].float().t() + packed["bf"].float())
    eps = y @ packed["wo"].float() + packed["bo"]
    eps = eps[:, None, :]
    if return_inter:
        return eps, inter
    return eps
