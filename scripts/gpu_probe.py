"""First-light GPU diagnostics (not a test): prints errors of every kernel family against the CPU oracle."""
import os, sys, time, traceback
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from norma_b200 import ffi, filters, synth
from oracle import mel_c
from oracle.whisper_oracle import Config, WhisperOracle, GreedyDecoder, special_tokens_for_vocab

def section(name):
    print(f"\n==== {name} ====", flush=True)

def run(name, fn):
    section(name)
    try:
        fn()
    except Exception:
        traceback.print_exc()
        print(f"[{name}] FAILED", flush=True)

cfgd = synth.model_config("test-micro")

def t_mel():
    for n_mel in (80, 128):
        c = dict(cfgd); c["num_mel_bins"] = n_mel
        ctx = ffi.Context(c, compute="f32", max_batch=2)
        f = filters.mel_filters(n_mel)
        ctx.set_mel_filters(f)
        for kind, n in (("gauss", 480000), ("uniform", 480000), ("bursts", 480000), ("gauss", 100000), ("gauss", 16000), ("gauss", 159), ("zeros", 4000), ("chirp", 480000)):
            pcm = synth.synth_pcm(kind, 0, n)
            ref = mel_c.pcm_to_mel(pcm, f)
            t = time.time(); got = ctx.pcm_to_mel(pcm); dt = time.time() - t
            print(f"n_mel={n_mel} {kind:8s} n={n:6d} shape {got.shape} vs {ref.shape} maxabs {np.abs(got-ref).max():.3e} ({dt*1e3:.1f} ms)", flush=True)
        ctx.close()

def t_gemm():
    rng = np.random.default_rng(0)
    for compute in ("f32", "bf16"):
        ctx = ffi.Context(cfgd, compute=compute, max_batch=1)
        for (M, N, K) in ((128, 256, 64), (128, 128, 64), (128, 256, 128), (256, 256, 256), (1500, 384, 384), (1500, 1152, 384), (300, 512, 1280), (3000, 1280, 240), (1000, 3840, 1280)):
            a = rng.standard_normal((M, K)).astype(np.float32); w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
            bias = rng.standard_normal(N).astype(np.float32)
            if compute == "bf16":
                a, w = ffi.bf16_round(a), ffi.bf16_round(w)
            ref = a.astype(np.float64) @ w.astype(np.float64).T + bias
            got = ctx.test_gemm(a, w, bias)
            err = np.abs(got - ref).max(); rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            print(f"{compute} gemm {M}x{N}x{K}: maxabs {err:.3e} rel {rel:.3e} nan={np.isnan(got).sum()}", flush=True)
            if compute == "bf16" and rel > 1e-3:
                bad = np.argwhere(np.abs(got - ref) > 1e-2)
                print("   first bad idx", bad[:5].tolist(), "n_bad", len(bad), "rows bad", np.unique(bad[:,0])[:10], "cols bad", np.unique(bad[:,1])[:10], flush=True)
        refg = None
        a = rng.standard_normal((256, 128)).astype(np.float32); w = (rng.standard_normal((256, 128)) * 0.1).astype(np.float32)
        if compute == "bf16": a, w = ffi.bf16_round(a), ffi.bf16_round(w)
        x = torch.from_numpy(a.astype(np.float64) @ w.astype(np.float64).T)
        refg = torch.nn.functional.gelu(x, approximate="tanh").numpy()
        got = ctx.test_gemm(a, w, None, gelu=True)
        print(f"{compute} gemm+gelu maxabs {np.abs(got-refg).max():.3e}", flush=True)
        ctx.close()

def t_attn():
    rng = np.random.default_rng(1)
    for compute in ("f32", "bf16"):
        ctx = ffi.Context(cfgd, compute=compute, max_batch=1)
        for (B, T, H) in ((1, 64, 2), (2, 200, 2), (1, 1500, 2)):
            d = H * 64
            qkv = rng.standard_normal((B * T, 3 * d)).astype(np.float32) * 0.5
            if compute == "bf16": qkv = ffi.bf16_round(qkv)
            q, k, v = [torch.from_numpy(qkv[:, i*d:(i+1)*d]).double().view(B, T, H, 64).transpose(1, 2) for i in range(3)]
            ref = (torch.softmax(q @ k.transpose(2, 3), -1) @ v).transpose(1, 2).reshape(B * T, d).numpy()
            os.environ["NB200_ATTN"] = "simt"
            got = ctx.test_attention(qkv, B, T, H)
            print(f"{compute} attn simt B{B} T{T} H{H}: maxabs {np.abs(got-ref).max():.3e}", flush=True)
        ctx.close()

def t_encoder():
    for name in ("test-micro", "tiny.en"):
        c = synth.model_config(name)
        w = synth.synth_weights(c, seed=1)
        orc = WhisperOracle(Config(**c), w)
        pcm = np.stack([synth.synth_pcm("gauss", 0), synth.synth_pcm("uniform", 1)])
        f = filters.mel_filters(c["num_mel_bins"])
        mel = np.stack([mel_c.pcm_to_mel(p, f)[:, :3000] for p in pcm])
        t = time.time(); ref, stages = orc.encoder_forward(torch.from_numpy(mel), return_stages=True); ref = ref.numpy(); print(f"oracle {name} {time.time()-t:.2f}s")
        for compute in ("f32", "bf16"):
            ctx = ffi.Context(c, compute=compute, max_batch=2)
            ctx.set_mel_filters(f); ctx.load_weights(w)
            got = ctx.encoder_forward(mel)
            e = np.abs(got - ref).max(); rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            print(f"{name} {compute} encoder_forward(mel): maxabs {e:.3e} rel {rel:.3e} nan={np.isnan(got).sum()}", flush=True)
            got2 = ctx.transcode_batch(pcm)
            e = np.abs(got2 - ref).max(); rel = np.linalg.norm(got2 - ref) / np.linalg.norm(ref)
            print(f"{name} {compute} transcode_batch(pcm): maxabs {e:.3e} rel {rel:.3e}", flush=True)
            ctx.close()

def t_decoder():
    name = "test-micro"
    c = synth.model_config(name)
    w = synth.synth_weights(c, seed=1)
    cfg = Config(**c)
    orc = WhisperOracle(cfg, w)
    f = filters.mel_filters(c["num_mel_bins"])
    pcm = np.stack([synth.synth_pcm("gauss", 0), synth.synth_pcm("uniform", 1)])
    mel = np.stack([mel_c.pcm_to_mel(p, f)[:, :3000] for p in pcm])
    xa = orc.encoder_forward(torch.from_numpy(mel))
    st = special_tokens_for_vocab(cfg.vocab_size)
    toks = [st.sot, st.lang, st.task, st.ts_zero, 11, 22, 333]
    for compute in ("f32", "bf16"):
        ctx = ffi.Context(c, compute=compute, max_batch=2)
        ctx.set_mel_filters(f); ctx.load_weights(w)
        ctx.set_tokens(st.sot, st.eot, st.task, st.lang, st.no_speech, st.no_timestamps, st.ts_zero, st.ts_one)
        ctx.encoder_forward(mel, want_output=False)
        for wdw in (0, 1):
            ref = orc.decoder_forward(torch.tensor([toks]), xa[wdw:wdw+1], True)[0].numpy()
            got = ctx.decoder_forward(toks, True, window=wdw)
            print(f"{compute} decoder_forward window {wdw}: maxabs {np.abs(got-ref).max():.3e}", flush=True)
        lg_ref = orc.final_linear(torch.from_numpy(ref[-1:]))[0].numpy()
        lg = ctx.final_linear(ref[-1])
        print(f"{compute} final_linear maxabs {np.abs(lg-lg_ref).max():.3e}", flush=True)
        res = ctx.decode_greedy(2, max_new_tokens=12)
        for wdw in (0, 1):
            dr = GreedyDecoder(orc, st).decode(xa[wdw:wdw+1], max_steps=12)
            print(f"{compute} greedy window {wdw}: same={dr.tokens == res[wdw]['tokens']} lp {dr.avg_logprob:.5f} vs {res[wdw]['avg_logprob']:.5f} nsp {dr.no_speech_prob:.3e} vs {res[wdw]['no_speech_prob']:.3e} min margin {min(dr.margins):.2e}")
            if dr.tokens != res[wdw]['tokens']: print("   ref", dr.tokens, "\n   got", res[wdw]['tokens'])
        ctx.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["mel", "gemm", "attn", "encoder", "decoder"]
    n = ffi.C.c_int(); print("device_count", ffi.load_library().nb200_device_count(ffi.C.byref(n)), n.value)
    for k in which:
        run(k, globals()["t_" + k])
