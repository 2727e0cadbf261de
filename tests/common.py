"""Shared fixtures for the parity tests: deterministic synthetic voices + oracle + product models."""
from __future__ import annotations

import functools

import numpy as np

from larynx_b200 import synth

GLOW_SEED = 1234
HIFI_SEED = 4321

# the five ljspeech fixture rows of the reference
# (/root/reference/local/en-us/ljspeech-glow_tts/samples/test_phonemes.csv, 111/28/47/51/61 ids)
LJSPEECH_FIXTURE_IDS = {
    "it_took_me_quite_a_long_time_to_develop_a_voice": "3 4 39 23 3 23 4 42 16 3 18 4 14 3 16 27 4 6 23 3 35 3 17 4 33 31 3 23 4 6 18 3 23 4 25 3 9 39 26 4 37 17 35 21 3 35 3 26 4 34 22 3 1 3 35 19 9 3 19 4 7 3 30 4 29 23 3 4 6 3 13 4 29 26 3 4 39 23 3 4 6 18 3 19 4 32 23 3 38 4 20 39 31 3 23 4 25 3 8 4 14 3 22 4 6 17 35 19 23 3 2",
    "be_a_voice_not_an_echo": "3 8 4 14 3 35 3 26 4 34 22 3 1 3 19 4 32 23 3 35 19 3 4 37 16 20 3 2",
}


def fixture_ids(name: str) -> np.ndarray:
    return np.array([int(x) for x in LJSPEECH_FIXTURE_IDS[name].split()], dtype=np.int64)


@functools.lru_cache(maxsize=None)
def glow_cfg(num_symbols: int = 46):
    return synth.glow_config(num_symbols)


@functools.lru_cache(maxsize=None)
def glow_sd_np(num_symbols: int = 46, seed: int = GLOW_SEED):
    return synth.glow_state_dict(glow_cfg(num_symbols), seed)


@functools.lru_cache(maxsize=None)
def hifi_cfg(name: str = "universal_large"):
    return synth.hifigan_config(name)


@functools.lru_cache(maxsize=None)
def hifi_sd_np(name: str = "universal_large", seed: int = HIFI_SEED):
    return synth.hifigan_state_dict(hifi_cfg(name), seed)


@functools.lru_cache(maxsize=None)
def oracle_glow_sd(num_symbols: int = 46, seed: int = GLOW_SEED):
    from oracle import port

    return port.to_torch_sd(glow_sd_np(num_symbols, seed))


@functools.lru_cache(maxsize=None)
def oracle_hifi_sd(name: str = "universal_large", seed: int = HIFI_SEED):
    from oracle import port

    return port.to_torch_sd(hifi_sd_np(name, seed))


@functools.lru_cache(maxsize=None)
def product_glow(num_symbols: int = 46, seed: int = GLOW_SEED, device: int = 0):
    from larynx_b200 import pack
    from larynx_b200.glow_tts import GlowTextToSpeech

    cfg = glow_cfg(num_symbols)
    return GlowTextToSpeech(None, device=device, blob=pack.pack_glow(glow_sd_np(num_symbols, seed), cfg), model_config=cfg)


@functools.lru_cache(maxsize=None)
def product_hifi(name: str = "universal_large", seed: int = HIFI_SEED, device: int = 0):
    from larynx_b200 import pack
    from larynx_b200.hifi_gan import HiFiGanVocoder

    cfg = hifi_cfg(name)
    return HiFiGanVocoder(None, device=device, blob=pack.pack_hifigan(hifi_sd_np(name, seed), cfg), model_config=cfg)
