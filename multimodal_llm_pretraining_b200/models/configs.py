"""Architecture hyper-parameters of the in-scope model families, hard-coded.

The reference pulls them from the HF hub (`GPTNeoXConfig.from_pretrained(f"EleutherAI/{model}")`,
src/models/pythia.py:18-21; `RobertaConfig.from_pretrained("roberta-large")`, src/models/roberta.py:16); there is no
network here, so the published values are entered by hand (parameter counts verified in
tests/test_registry.py)."""
from __future__ import annotations

from types import SimpleNamespace

# name: (layers, hidden, heads, vocab)
_PYTHIA = {
    "pythia-14m": (6, 128, 4, 50304),
    "pythia-31m": (6, 256, 8, 50304),
    "pythia-70m": (6, 512, 8, 50304),
    "pythia-160m": (12, 768, 12, 50304),
    "pythia-410m": (24, 1024, 16, 50304),
    "pythia-1b": (16, 2048, 8, 50304),
    "pythia-1.4b": (24, 2048, 16, 50304),
    "pythia-2.8b": (32, 2560, 32, 50304),
    "pythia-6.9b": (32, 4096, 32, 50432),
    "pythia-12b": (36, 5120, 40, 50688),
}

PYTHIA_PARAM_COUNTS = {
    "pythia-14m": 14_067_712, "pythia-31m": 30_494_720, "pythia-70m": 70_426_624, "pythia-160m": 162_322_944,
    "pythia-410m": 405_334_016, "pythia-1b": 1_011_781_632, "pythia-1.4b": 1_414_647_808,
    "pythia-2.8b": 2_775_208_960, "pythia-6.9b": 6_857_302_016, "pythia-12b": 11_846_072_320,
}
ROBERTA_LARGE_PARAM_COUNT = 355_412_057


def pythia_config_dict(model_type: str) -> dict:
    L, h, nh, V = _PYTHIA[model_type]
    return dict(vocab_size=V, hidden_size=h, num_hidden_layers=L, num_attention_heads=nh, intermediate_size=4 * h,
                max_position_embeddings=2048, rotary_pct=0.25, rotary_emb_base=10000, layer_norm_eps=1e-5,
                use_parallel_residual=True, hidden_act="gelu", attention_bias=True, hidden_dropout=0.0,
                attention_dropout=0.0, tie_word_embeddings=False, initializer_range=0.02, bos_token_id=0, eos_token_id=0)


def roberta_large_config_dict() -> dict:
    return dict(vocab_size=50265, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5, hidden_dropout_prob=0.1,
                attention_probs_dropout_prob=0.1, pad_token_id=1, bos_token_id=0, eos_token_id=2, hidden_act="gelu",
                initializer_range=0.02, tie_word_embeddings=True)


def as_namespace(d: dict) -> SimpleNamespace:
    ns = SimpleNamespace(**d)
    ns.to_dict = lambda: dict(d)  # HF Trainer / logging call config.to_dict()
    return ns


def neox_param_count(d: dict) -> int:
    h, V, L, I = d["hidden_size"], d["vocab_size"], d["num_hidden_layers"], d["intermediate_size"]
    per_layer = 4 * h + (3 * h * h + 3 * h) + (h * h + h) + (I * h + I) + (h * I + h)
    return 2 * V * h + L * per_layer + 2 * h


def neox_linear_weight_count(d: dict) -> int:
    """W_lin of SURVEY.md §8d: all 2-D Linear weights incl. the LM head, excl. the input embedding."""
    h, V, L, I = d["hidden_size"], d["vocab_size"], d["num_hidden_layers"], d["intermediate_size"]
    return L * (3 * h * h + h * h + 2 * I * h) + V * h


def neox_train_flops_per_sequence(d: dict, S: int) -> int:
    """F(S) = 6 S W_lin + 12 L h S^2 — equals torch FlopCounterMode on fwd+bwd, the reference's FLOP definition
    (src/benchmarking/flops.py:28-36)."""
    return 6 * S * neox_linear_weight_count(d) + 12 * d["num_hidden_layers"] * d["hidden_size"] * S * S


def roberta_linear_weight_count(d: dict) -> int:
    """W_lin for RoBERTa: encoder Linear weights + lm_head.dense + the tied decoder (counted once as a GEMM)."""
    h, V, L, I = d["hidden_size"], d["vocab_size"], d["num_hidden_layers"], d["intermediate_size"]
    return L * (4 * h * h + 2 * I * h) + h * h + V * h


def roberta_train_flops_per_sequence(d: dict, S: int) -> int:
    """F(S) = 6 S W_lin + 12 L h S^2 (bidirectional attention: no causal discount applies)."""
    return 6 * S * roberta_linear_weight_count(d) + 12 * d["num_hidden_layers"] * d["hidden_size"] * S * S
