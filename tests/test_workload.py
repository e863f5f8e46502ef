"""The algorithmic work figures bench.py quotes its rooflines against (SURVEY.md §8 d; norma_b200/workload.py), checked against hand counts
for the distil-large-v3 shape and against an op-by-op count of the oracle's graph for a small shape.  CPU only."""
import pytest

from norma_b200 import synth, workload as wl


def test_distil_large_v3_hand_counts():
    c = synth.model_config("distil-large-v3")
    d, L, nm, T = 1280, 32, 128, 1500
    assert (c["d_model"], c["encoder_layers"], c["num_mel_bins"]) == (d, L, nm)
    conv = 2.0 * 3000 * d * 3 * nm + 2.0 * T * d * 3 * d
    layer_gemm = 2.0 * T * d * (3 * d) + 2.0 * T * d * d + 2.0 * T * d * (4 * d) * 2  # qkv, out, fc1 + fc2
    layer_attn = 2.0 * T * T * d * 2  # q.k^T and p.v over all heads
    assert wl.gemm_flops(c) == pytest.approx(conv + L * layer_gemm)
    assert wl.attention_flops(c) == pytest.approx(L * layer_attn)
    assert wl.encoder_flops(c) == pytest.approx(conv + L * (layer_gemm + layer_attn))
    assert wl.encoder_flops(c) / 1e9 == pytest.approx(2273.77, rel=1e-4)  # the per-window figure DESIGN.md and the VERDICT quote
    assert wl.mel_bytes(c) == 4 * 480_000 + 4 * nm * 3000
    # one decoder step streams every decoder weight once: per layer q,k,v,o + cross q,o (6 d^2) + fc1, fc2 (8 d^2) = 14 d^2, plus the tied
    # embedding (V d), plus the cross-attention K/V of each window (L x 1500 x 2d)
    V, Ld = c["vocab_size"], c["decoder_layers"]
    assert wl.decode_bytes_per_step(c, 1) == 2 * (Ld * 14 * d * d + V * d) + Ld * T * 2 * d * 2
    assert wl.decode_bytes_per_step(c, 8) - wl.decode_bytes_per_step(c, 1) == 7 * Ld * T * 2 * d * 2


def test_gemm_bytes_counts_the_folded_layernorm():
    c = synth.model_config("distil-large-v3")
    d, B = c["d_model"], 25
    M = B * wl.T_ENC
    plain, folded = wl.gemm_bytes(c, B, 2, ln_folded=False), wl.gemm_bytes(c, B, 2, ln_folded=True)
    # per layer the fold adds two bf16 copies of the residual rows (out-proj, fc2) and four passes over 2 x d / 256 float2 slots per row
    extra = 2 * (2.0 * M * d) + 4 * (8.0 * 2 * (d // 256) * M)
    assert 4 * (folded - plain) == pytest.approx(extra)
    # and removes two LayerNorm kernels per layer from the step, each reading f32 and writing bf16 rows: more than it adds
    removed = 2 * M * wl.layernorm_bytes_per_row(d, 2)
    assert removed > extra
    # out-proj: A (bf16) + W + f32 residual in and out = 480 MB at 25 windows, the figure profiles/r1e_gemm_full_summary.md measured as DRAM traffic
    out_proj = 2 * (M * d + d * d) + 8.0 * M * d
    assert out_proj / 1e6 == pytest.approx(483.3, rel=1e-3)


def test_flops_match_an_op_count_of_the_graph():
    """encoder_flops against a literal walk over the layers of a small config (every matmul of WhisperOracle.encoder_forward)."""
    c = synth.model_config("tiny.en")
    d, L, nm, T, H = c["d_model"], c["encoder_layers"], c["num_mel_bins"], 1500, c["encoder_attention_heads"]
    total = 2 * 3000 * (3 * nm) * d  # conv1 as a GEMM over 3 taps
    total += 2 * T * (3 * d) * d     # conv2, stride 2
    for _ in range(L):
        total += 3 * (2 * T * d * d)             # q, k, v projections
        total += H * (2 * T * T * (d // H)) * 2  # scores and context per head
        total += 2 * T * d * d                   # out projection
        total += 2 * (2 * T * d * 4 * d)         # fc1, fc2
    assert wl.encoder_flops(c) == pytest.approx(float(total))


def test_sharding_helpers():
    assert wl.windows_of_rank(1, 4, 10) == [1, 5, 9]
    assert wl.window_ids(2, 8, 25) == list(range(50, 75))
    assert wl.window_ids(3, 8, 25, total=120) == list(range(3, 120, 8))
    assert wl.windows_per_step(8, 25) == 200 and wl.windows_per_step(8, 25, 120) == 120
    assert sorted(w for r in range(8) for w in wl.windows_of_rank(r, 8, 120)) == list(range(120))
    with pytest.raises(ValueError):
        wl.windows_of_rank(4, 4, 10)
