"""Fake data backend with the reference's constructor (tssep/data.py:11-152).

Deterministic 8-speaker single-channel sinusoid mixtures seeded by
``np.random.RandomState(seed)``; used by the shipped toy configs and by the
throughput benchmark (``num_samples`` generalises the reference's fixed 5 s).
"""
from __future__ import annotations

import dataclasses

import numpy as np


@dataclasses.dataclass
class DummyReader:
    train_dataset_name: str = "train"
    validate_dataset_name: str = "validate"
    domain_adaptation_src_dataset_name: str = "validate"
    eval_dataset_name: str = "eval"
    sample_rate: int = 16000
    aux_size: int = 100
    train_examples: int = 10

    num_speakers = 8

    def _get_vad(self, num_samples, num_speakers):
        vad = np.zeros((num_speakers, num_samples), dtype=bool)
        start = 0
        for i in range(num_speakers):
            end = num_samples * (i + 2) // (num_speakers + 1)
            vad[i, start:end] = True
            start = end - (end - start) // 2
        return vad

    def get_example(self, seed, num_samples=None, dataset_name="validate", with_targets=True):
        K = self.num_speakers
        if num_samples is None:
            num_samples = self.sample_rate * 5
        rng = np.random.RandomState(seed)
        frequency = rng.randint(100, 7000, size=(3, K))
        time = np.arange(num_samples) / self.sample_rate
        vad = self._get_vad(num_samples, K)
        early = np.empty((K, num_samples), dtype=np.float32)
        for k in range(K):
            acc = np.sin(2 * np.pi * frequency[0, k] * time)
            acc += np.sin(2 * np.pi * frequency[1, k] * time)
            acc += np.sin(2 * np.pi * frequency[2, k] * time)
            early[k] = acc
        early *= vad
        noise = rng.rand(1, num_samples).astype(np.float32)
        observation = early.sum(axis=0, keepdims=True) + noise
        aux = np.zeros((K, self.aux_size), dtype=np.float32)
        for spk, fs in enumerate(frequency.T):
            for f in fs:
                f = (f * self.aux_size) // 7001
                aux[spk, f:f + 2] = 1
        ex = {
            "example_id": f"dummy_id_{seed}",
            "num_samples": num_samples,
            "audio_data": {"observation": observation, "vad": vad},
            "auxInput": aux,
            "dataset": dataset_name,
        }
        if with_targets:
            ex["audio_data"]["speaker_reverberation_early_ch0"] = early
        return ex

    def __call__(self, dataset_name, pre_load_apply=None, load_keys=("observation",)):
        n = self.train_examples if "train" in dataset_name else 4
        with_targets = "speaker_reverberation_early_ch0" in load_keys
        ds = [self.get_example(i, dataset_name=dataset_name, with_targets=with_targets) for i in range(n)]
        return pre_load_apply(ds) if pre_load_apply is not None else ds

    class data_hooks:
        @staticmethod
        def pre_net(ex):
            return ex
