"""The model oracle against an independent implementation of the same graph (HF transformers WhisperModel with the
conv-stem GELU patched to the tanh form candle uses), its own invariants, and the committed golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden
from norma_b200 import filters, synth
from oracle import mel_c
from oracle.whisper_oracle import Config, GreedyDecoder, WhisperOracle, sinusoids, sinusoids_numpy, special_tokens_for_vocab


@pytest.fixture(scope="module")
def micro():
    c = synth.model_config("test-micro")
    w = synth.synth_weights(c, seed=1)
    return c, w, WhisperOracle(Config(**c), w)


def test_hf_names_cover_whisper_state_dict(micro):
    transformers = pytest.importorskip("transformers")
    c, w, _ = micro
    hc = transformers.WhisperConfig(vocab_size=c["vocab_size"], num_mel_bins=c["num_mel_bins"], encoder_layers=c["encoder_layers"],
                                    encoder_attention_heads=c["encoder_attention_heads"], decoder_layers=c["decoder_layers"],
                                    decoder_attention_heads=c["decoder_attention_heads"], d_model=c["d_model"], encoder_ffn_dim=4 * c["d_model"],
                                    decoder_ffn_dim=4 * c["d_model"])
    sd = transformers.WhisperModel(hc).state_dict()
    ours = {k[len("model."):] for k in w}
    assert set(sd) - ours == {"encoder.embed_positions.weight"}  # ignored: candle recomputes sinusoids
    assert ours - set(sd) == set()
    for k, v in sd.items():
        if "model." + k in w:
            assert tuple(v.shape) == tuple(w["model." + k].shape), k


def test_encoder_decoder_match_hf(micro):
    transformers = pytest.importorskip("transformers")
    import transformers.models.whisper.modeling_whisper as mw

    c, w, orc = micro
    hc = transformers.WhisperConfig(vocab_size=c["vocab_size"], num_mel_bins=c["num_mel_bins"], encoder_layers=c["encoder_layers"],
                                    encoder_attention_heads=c["encoder_attention_heads"], decoder_layers=c["decoder_layers"],
                                    decoder_attention_heads=c["decoder_attention_heads"], d_model=c["d_model"], encoder_ffn_dim=4 * c["d_model"],
                                    decoder_ffn_dim=4 * c["d_model"], activation_function="gelu_pytorch_tanh", max_source_positions=1500,
                                    max_target_positions=448, attn_implementation="eager")
    hm = transformers.WhisperModel(hc).eval()
    sd = hm.state_dict()
    for k in sd:
        if "model." + k in w:
            sd[k] = w["model." + k].clone()
    sd["encoder.embed_positions.weight"] = sinusoids(1500, c["d_model"])  # same table on both sides
    hm.load_state_dict(sd)
    orig = torch.nn.functional.gelu
    mw.nn.functional.gelu = lambda x, approximate="none": orig(x, approximate="tanh")  # HF stem uses erf-GELU
    try:
        mel = torch.randn(1, c["num_mel_bins"], 3000, generator=torch.Generator().manual_seed(0)) * 0.5
        with torch.no_grad():
            he = hm.encoder(mel).last_hidden_state
        oe = orc.encoder_forward(mel)
        assert (he - oe).abs().max() < 2e-5
        toks = torch.tensor([[50257, 50258, 50358, 50363, 11, 22]])
        with torch.no_grad():
            hd = hm.decoder(input_ids=toks, encoder_hidden_states=he).last_hidden_state
        od = orc.decoder_forward(toks, oe, True)
        assert (hd - od).abs().max() < 2e-5
    finally:
        mw.nn.functional.gelu = orig


def test_sinusoids_layout_and_libm_vs_numpy():
    s = sinusoids(1500, 384)
    assert s.shape == (1500, 384)
    assert torch.all(s[0, :192] == 0) and torch.all(s[0, 192:] == 1)  # [sin | cos] halves at t = 0
    assert (s - sinusoids_numpy(1500, 384)).abs().max() < 3e-4  # one ulp of inv_timescale at t ~ 1500 rad


def test_cross_kv_cache_semantics(micro):
    c, w, orc = micro
    g = torch.Generator().manual_seed(3)
    xa1, xa2 = torch.randn(1, 1500, c["d_model"], generator=g), torch.randn(1, 1500, c["d_model"], generator=g)
    toks = torch.tensor([[50257, 50258, 50358]])
    a = orc.decoder_forward(toks, xa1, True)
    b = orc.decoder_forward(toks, xa2, False)  # flush = false: cached K/V of xa1 are reused
    assert torch.equal(a, b)
    cc = orc.decoder_forward(toks, xa2, True)
    assert not torch.allclose(a, cc)
    orc.reset_kv_cache()
    assert orc.cross_kv is None


def test_causality_prefix_invariance(micro):
    """A self-attention KV cache is numerically equivalent to the reference's full recompute: position i of the
    decoder output depends only on tokens[:i+1]."""
    c, w, orc = micro
    xa = torch.randn(1, 1500, c["d_model"], generator=torch.Generator().manual_seed(4))
    toks = [50257, 50258, 50358, 50363, 7, 8, 9]
    full = orc.decoder_forward(torch.tensor([toks]), xa, True)
    for n in (1, 3, 5):
        part = orc.decoder_forward(torch.tensor([toks[:n]]), xa, False)
        assert (part[0] - full[0, :n]).abs().max() < 2e-6


def test_greedy_rules(micro):
    c, w, orc = micro
    st = special_tokens_for_vocab(c["vocab_size"])
    xa = torch.randn(1, 1500, c["d_model"], generator=torch.Generator().manual_seed(5)) * 0.5
    dr = GreedyDecoder(orc, st).decode(xa, max_steps=10)
    assert dr.tokens[:3] == [st.sot, st.lang, st.task]
    assert st.ts_zero <= dr.tokens[3] <= st.ts_one  # first sampled token is forced into [<|0.00|>, <|1.00|>]
    assert dr.tokens[-1] == st.eot
    assert st.no_timestamps not in dr.tokens[3:]  # suppressed together with Config::suppress_tokens
    ts = [t for t in dr.tokens[3:-1] if t > st.no_timestamps]
    assert ts == sorted(ts)  # timestamps never go backwards


def test_golden_encoder_tiny_en():
    g = golden("enc_tiny_en.npz")
    c = synth.model_config("tiny.en")
    orc = WhisperOracle(Config(**c), synth.synth_weights(c, seed=1, decoder=False))
    mel = mel_c.pcm_to_mel(synth.synth_pcm("gauss", 0), filters.mel_filters(80))[:, :3000]
    y = orc.encoder_forward(torch.from_numpy(mel[None]))[0].numpy()
    assert np.abs(y[g["rows"]] - g["values"]).max() < 2e-5
    assert abs(np.linalg.norm(y.astype(np.float64)) - float(g["fro"])) < 1e-3 * float(g["fro"])
