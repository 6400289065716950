"""Generates tests/golden/attention_golden.npz (run in the BUILD container only; needs /root/reference).

    python tests/golden/make_golden_attention.py

The reference's OWN TransformerDecoderLayer (model/transformer_detr.py, imported unmodified; it needs only torch and
model.helper) with use_rel=True, random weights, run through forward_pre_rel on CPU.  Forward hooks capture what the
cross-attention block (:443-454) sees and produces: norm2's output (= tgt2), the input of attn_mlp (checks the
x = tgt2 - memory + relative_pos bookkeeping) and the output of out_mlp.  Cases: ragged context counts (not a
multiple of the kernel's 128-context tile), batch 2, embeddings in [-1, 1] like the Fourier features."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def load_layer_class():
    spec = importlib.util.spec_from_file_location("_ref_detr", "/root/reference/model/transformer_detr.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.TransformerDecoderLayer


def main():
    Layer = load_layer_class()
    out = {}
    cases = [(8, 200, 2), (4, 128, 1), (3, 333, 1)]
    for i, (Q, C, B) in enumerate(cases):
        torch.manual_seed(100 + i)
        layer = Layer(64, nhead=4, dim_feedforward=256, dropout=0.0, use_rel=True).eval()
        tgt = torch.randn(Q, B, 64)
        memory = torch.randn(C, B, 64)
        pos = torch.randn(C, B, 64)
        qpos = torch.randn(Q, B, 64)
        rel = torch.rand(Q, C, B, 64) * 2 - 1
        cap = {}
        hooks = [layer.norm2.register_forward_hook(lambda m, a, o: cap.__setitem__("tgt2", o.detach().clone())),
                 layer.attn_mlp.register_forward_hook(lambda m, a, o: cap.__setitem__("x", a[0].detach().clone())),
                 layer.out_mlp.register_forward_hook(lambda m, a, o: cap.__setitem__("out", o.detach().clone()))]
        with torch.no_grad():
            layer.forward_pre_rel(tgt, memory, pos=pos, query_pos=qpos, relative_pos=rel)
        for h in hooks:
            h.remove()
        assert torch.equal(cap["x"], cap["tgt2"][:, None] - memory[None] + rel)
        out.update({"c%d_tgt2" % i: cap["tgt2"].numpy(), "c%d_memory" % i: memory.numpy(), "c%d_rel" % i: rel.numpy(),
                    "c%d_out" % i: cap["out"].numpy(),
                    "c%d_w1" % i: layer.attn_mlp[0].weight.detach().numpy(), "c%d_b1" % i: layer.attn_mlp[0].bias.detach().numpy(),
                    "c%d_w2" % i: layer.attn_mlp[2].weight.detach().numpy(), "c%d_b2" % i: layer.attn_mlp[2].bias.detach().numpy(),
                    "c%d_wv" % i: layer.v_mlp[0].weight.detach().numpy(), "c%d_bv" % i: layer.v_mlp[0].bias.detach().numpy(),
                    "c%d_wo" % i: layer.out_mlp[0].weight.detach().numpy(), "c%d_bo" % i: layer.out_mlp[0].bias.detach().numpy()})
    out["n"] = np.array(len(cases))
    path = os.path.join(HERE, "attention_golden.npz")
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in out.items()})
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
