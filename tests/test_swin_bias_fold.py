"""Host logic behind ysod_swin64_tc's bias contract (include/ysod.h, engine.py `swin`): nn.MultiheadAttention(x) is unchanged when the key
third of in_proj_bias is dropped (a bias on every key shifts all scores of a query equally: softmax-invariant) and the value third is
moved into out_proj's bias (rows of softmax sum to 1): bo' = bo + Wo bv. Reference: blocks_transformer.py:98-116 (WindowAttention.attn)."""
import torch


def test_key_bias_drops_and_value_bias_folds_into_out_proj():
    torch.manual_seed(3)
    E, heads, T, B = 64, 2, 49, 5
    mha = torch.nn.MultiheadAttention(E, heads, batch_first=True).double().eval()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.5)
        mha.out_proj.bias.normal_(0, 0.5)
    x = torch.randn(B, T, E, dtype=torch.double)
    with torch.no_grad():
        ref = mha(x, x, x, need_weights=False)[0]
        folded = torch.nn.MultiheadAttention(E, heads, batch_first=True).double().eval()
        folded.load_state_dict(mha.state_dict())
        bv = mha.in_proj_bias[2 * E:].clone()
        folded.in_proj_bias[E:] = 0.0                                   # key and value thirds
        folded.out_proj.bias += mha.out_proj.weight @ bv                # bo' = bo + Wo bv
        got = folded(x, x, x, need_weights=False)[0]
    assert float((got - ref).abs().max()) < 1e-12 * max(1.0, float(ref.abs().max())) * 1e2
