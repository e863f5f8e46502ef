"""Host-side arithmetic of the benchmarked workloads: how windows are sharded over ranks / contexts (SURVEY.md §8 e: window w -> GPU
w mod G, no collective) and the ALGORITHMIC work figures the rooflines are quoted against (SURVEY.md §8 d).  Pure Python, importable
without a GPU: bench.py uses these, tests/test_sharding_gloo.py and tests/test_workload.py check them."""
from __future__ import annotations

from typing import Dict, List

WINDOW_S = 30.0
N_SAMPLES = 480_000
N_FRAMES = 3000
T_ENC = 1500


# ---------------------------------------------------------------------------------------------------- sharding
def windows_of_rank(rank: int, world: int, total: int) -> List[int]:
    """Strong-scaling mode (BASELINE config 3: a fixed pool of `total` windows): window w belongs to rank w mod world."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"windows_of_rank(rank={rank}, world={world}, total={total})")
    return list(range(rank, total, world))


def window_ids(rank: int, world: int, per_rank: int, total: int = 0) -> List[int]:
    """Global ids (= PCM seeds) of the windows a rank processes per step: `total` > 0 selects the strong-scaling pool, otherwise every
    rank owns `per_rank` windows of its own (weak scaling: rank r holds r * per_rank .. r * per_rank + per_rank - 1)."""
    if total:
        return windows_of_rank(rank, world, total)
    return [rank * per_rank + i for i in range(per_rank)]


def windows_per_step(world: int, per_rank: int, total: int = 0) -> int:
    return total if total else world * per_rank


# ---------------------------------------------------------------------------------------------------- algorithmic work
def encoder_flops(c: Dict[str, int]) -> float:
    """2MNK FLOPs of one window: conv1 + conv2 + L x (qkv/out projections + attention + MLP)."""
    d, L, nm = c["d_model"], c["encoder_layers"], c["num_mel_bins"]
    return 2.0 * N_FRAMES * d * 3 * nm + 2.0 * T_ENC * d * 3 * d + L * (8.0 * T_ENC * d * d + 4.0 * T_ENC * T_ENC * d + 16.0 * T_ENC * d * d)


def gemm_flops(c: Dict[str, int]) -> float:
    d, L, nm = c["d_model"], c["encoder_layers"], c["num_mel_bins"]
    return 2.0 * N_FRAMES * d * 3 * nm + 2.0 * T_ENC * d * 3 * d + L * (8.0 * T_ENC * d * d + 16.0 * T_ENC * d * d)


def attention_flops(c: Dict[str, int]) -> float:
    return c["encoder_layers"] * 4.0 * T_ENC * T_ENC * c["d_model"]


def mel_bytes(c: Dict[str, int]) -> float:
    """PCM in (f32) + log-mel out (f32), per window."""
    return 4.0 * N_SAMPLES + 4.0 * c["num_mel_bins"] * N_FRAMES


def layernorm_bytes_per_row(d: int, out_bytes: int = 2) -> float:
    return (4.0 + out_bytes) * d


def gemm_bytes(c: Dict[str, int], B: int, es: int = 2, ln_folded: bool = True) -> float:
    """Algorithmic HBM bytes of one layer's four GEMMs at B windows (operands once, f32 residual in and out), mean per launch.
    `ln_folded` (the bf16 build): out-proj and fc2 also write the bf16 copy of the residual rows and 2 x d / 256 float2 statistics slots per row,
    the QKV / fc1 GEMMs read the slots (the 2 x (4 + 2) d bytes per row of the two LayerNorm kernels they replace are gone from the step)."""
    d, M = c["d_model"], B * T_ENC
    slots = 8.0 * 2 * max(1, d // 256) * M if ln_folded else 0.0  # 256-wide pair tiles x 2 epilogue warp groups
    copy = 2.0 * M * d if ln_folded else 0.0
    qkv = es * (M * d + 3 * d * d + M * 3 * d) + slots
    out = es * (M * d + d * d) + 8.0 * M * d + copy + slots
    fc1 = es * (M * d + 4 * d * d + M * 4 * d) + slots
    fc2 = es * (M * 4 * d + 4 * d * d) + 8.0 * M * d + copy + slots
    return (qkv + out + fc1 + fc2) / 4.0


def decode_bytes_per_step(c: Dict[str, int], B: int = 1, es: int = 2) -> float:
    """Weight streaming of one decoder step (2 (L 14 d^2 + V d) bytes in bf16) plus, per window, the cross-attention K/V it re-reads."""
    d, L, V = c["d_model"], c["decoder_layers"], c["vocab_size"]
    return es * (L * 14.0 * d * d + V * d) + B * L * T_ENC * 2.0 * d * es
