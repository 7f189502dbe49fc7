"""CUDA-graph replay of the consumer around the hot path.

One APGD iteration is: image update -> ``model(x_adv)`` -> fused loss -> ``d logits / d x_adv``
-> bookkeeping (semseg/attacker.py:385-551).  The robseg kernels are a handful of launches; the
model's forward and input-gradient backward are several hundred small ones, and below ~8 images
per batch the GPU waits for the host to issue them.  :class:`GraphedModel` captures the two
passes of a frozen, eval-mode model once for a fixed input shape and replays them; the attack
(``apgd_train`` / ``apgd_largereps``) uses ``model(x)`` and ``model.input_grad(dlogits)`` when the
model offers them.  The fused-loss kernel writes its
gradient straight into the graph's static ``gout`` buffer, so no logits-sized copy is added.

The parameters are constants of the attack: their ``requires_grad`` is switched off while the
backward is captured so the graph holds only the input-gradient path (what
``torch.autograd.grad(loss, [x_adv])`` computes in the reference, attacker.py:350,469).
"""
import torch


class GraphedModel:
    training = False  # apgd_train asserts ``not model.training`` (attacker.py:280)

    def __init__(self, model, example, warmup=3):
        if model.training:
            raise ValueError("GraphedModel captures an eval-mode model")
        if not example.is_cuda:
            raise RuntimeError("GraphedModel needs a CUDA example input (no CPU fallback)")
        self.model = model
        self.x = example.detach().float().contiguous().clone().requires_grad_(True)
        req = [p.requires_grad for p in model.parameters()]
        for p in model.parameters():
            p.requires_grad_(False)
        try:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # warm-up off the capture stream: cuDNN plans, lazy inits
                for _ in range(warmup):
                    out = model(self.x)
                    torch.autograd.grad(out, [self.x], grad_outputs=torch.zeros_like(out))
                del out
            cur.wait_stream(side)
            self.fwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.fwd):
                self.logits = model(self.x)
            self.gout = torch.zeros_like(self.logits)
            self.bwd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.bwd, pool=self.fwd.pool()):
                (self.gx,) = torch.autograd.grad(self.logits, [self.x], grad_outputs=self.gout)
        finally:
            for p, r in zip(model.parameters(), req):
                p.requires_grad_(r)

    def eval(self):
        return self

    def parameters(self):
        return self.model.parameters()

    def __call__(self, x):
        """logits of x, in the graph's static output buffer (overwritten by the next call)."""
        if x.shape != self.x.shape:
            raise ValueError(f"GraphedModel captured for input {tuple(self.x.shape)}, got {tuple(x.shape)}")
        with torch.no_grad():
            self.x.copy_(x)
        self.fwd.replay()
        return self.logits

    def input_grad(self, gout):
        """(d logits / d x)^T gout for the last forward; gout may BE ``self.gout`` (no copy)."""
        if gout.data_ptr() != self.gout.data_ptr():
            self.gout.copy_(gout)
        self.bwd.replay()
        return self.gx
