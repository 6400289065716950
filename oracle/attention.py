"""CPU restatement of the decoder's vector cross-attention over the relative-position embedding -- TEST
INFRASTRUCTURE ONLY.  Plain fp32 torch ops following model/transformer_detr.py:443-454 line by line (the MLPs are
defined at :384-396).  PINNED by tests/golden/attention_golden.npz: inputs, weights and the out_mlp output of the
reference's OWN TransformerDecoderLayer.forward_pre_rel, imported from /root/reference and run on CPU
(tests/golden/make_golden_attention.py; tests/test_oracle_golden.py asserts this function == the fixture)."""
import numpy as np
import torch
import torch.nn.functional as F


def rel_cross_attention(tgt2, memory, relative_pos, w):
    """tgt2 (Q,B,64) = norm2(tgt), memory (C,B,64), relative_pos (Q,C,B,64);
    w: dict with w1,b1,w2,b2 (attn_mlp), wv,bv (v_mlp), wo,bo (out_mlp) -> (Q,B,64) = out_mlp(sum_c attn * v2)"""
    n_queries, n_context, batch, channel = relative_pos.shape  # :443
    tgt2_expand = tgt2[:, None, :, :].repeat(1, n_context, 1, 1)  # :444
    memory_expand = memory[None, :, :, :].repeat(n_queries, 1, 1, 1)  # :446
    x = tgt2_expand - memory_expand + relative_pos
    sim = F.linear(F.relu(F.linear(x, w["w1"], w["b1"])), w["w2"], w["b2"])  # :448 attn_mlp
    attn = F.softmax(sim / np.sqrt(sim.shape[-1]), dim=1)  # :449
    v2 = F.linear(memory_expand + relative_pos, w["wv"], w["bv"])  # :451 v_mlp
    out = torch.einsum("qcbf,qcbf->qbf", attn, v2)  # :452
    return F.relu(F.linear(out, w["wo"], w["bo"]))  # :453 out_mlp
