"""Length-bucketed micro-batch scheduler for the fused ids -> PCM path.

The reference submits one `_sentence_task` per sentence to a ThreadPoolExecutor and yields the futures in submission
order (larynx/__init__.py:146-157, 187-190).  On one B200 a single sentence does not fill the device (the GlowTTS half
is a chain of small launches), so this scheduler

* groups the sentences of a request into micro-batches of similar phoneme count (`dispatch.bucket_by_length`), longest
  first, bounded by `max_batch` items and `max_cost` padded phonemes;
* runs the micro-batches on a few host threads, each with its own call context (stream + workspaces), through
  `host.synthesize_batch` (b200tts_glow_encode + b200tts_tts_finish: no host round trip between the two passes);
* hands the audio back in submission order, either all at once (`synthesize`) or as an iterator that yields sentence i
  as soon as sentences 0..i are done (`stream`), which is what the reference's generator does.

Measured on one B200 (tools/throughput.py, profiles/r1_tcgen05_summary.md): ragged U{64..256}-phoneme sentences,
micro-batches of 16: 248 k mel frames/s in arrival order, 302 k length-bucketed.
"""
from __future__ import annotations

import threading
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import dispatch

SynthFn = Callable[[List[np.ndarray], int], List[np.ndarray]]


class BatchScheduler:
    def __init__(self, tts_model=None, vocoder_model=None, glue=None, max_batch: int = 16, max_cost: Optional[int] = 4096,
                 workers: int = 2, noise_scale: float = 0.667, length_scale: float = 1.0, denoiser_strength: float = 0.0,
                 synth: Optional[SynthFn] = None):
        """`synth(rows, seed) -> list of int16 arrays` replaces the device path (tests); by default it is
        `host.synthesize_batch` on the given models."""
        assert max_batch >= 1 and workers >= 1
        self.max_batch, self.max_cost, self.workers = int(max_batch), max_cost, int(workers)
        if synth is None:
            assert tts_model is not None and vocoder_model is not None and glue is not None
            from . import host

            def synth(rows: List[np.ndarray], seed: int) -> List[np.ndarray]:
                return host.synthesize_batch(tts_model, vocoder_model, glue, rows, noise_scale=noise_scale,
                                             length_scale=length_scale, seed=seed, denoiser_strength=denoiser_strength)["audio"]

        self._synth = synth

    # ------------------------------------------------------------------------------------------------------------
    def plan(self, id_rows: Sequence[np.ndarray]) -> List[List[int]]:
        """Micro-batches (lists of sentence indices), longest first; empty sentences are left out."""
        lens = [len(r) for r in id_rows]
        keep = [i for i, n in enumerate(lens) if n > 0]
        buckets = dispatch.bucket_by_length([lens[i] for i in keep], self.max_batch, self.max_cost)
        return [[keep[j] for j in b] for b in buckets][::-1]

    def _run(self, id_rows: Sequence[np.ndarray], seed: int, on_done: Callable[[int, np.ndarray], None]):
        rows = [np.asarray(r, dtype=np.int64).reshape(-1) for r in id_rows]
        for i, r in enumerate(rows):
            if len(r) == 0:
                on_done(i, np.zeros(0, dtype=np.int16))
        plan = self.plan(rows)
        nxt = [0]
        lock = threading.Lock()
        errors: List[BaseException] = []

        def worker():
            while not errors:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= len(plan):
                    return
                idx = plan[k]
                try:
                    audio = self._synth([rows[i] for i in idx], seed + k)
                    assert len(audio) == len(idx)
                    for i, a in zip(idx, audio):
                        on_done(i, a)
                except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                    errors.append(e)

        threads = [threading.Thread(target=worker, daemon=True) for _ in range(min(self.workers, max(len(plan), 1)))]
        for t in threads:
            t.start()
        return threads, errors

    def synthesize(self, id_rows: Sequence[np.ndarray], seed: int = 0) -> List[np.ndarray]:
        """int16 audio of every sentence, in submission order."""
        out: List[Optional[np.ndarray]] = [None] * len(id_rows)

        def on_done(i: int, a: np.ndarray):
            out[i] = a

        threads, errors = self._run(id_rows, seed, on_done)
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        assert all(a is not None for a in out)
        return out  # type: ignore[return-value]

    def stream(self, id_rows: Sequence[np.ndarray], seed: int = 0) -> Iterator[Tuple[int, np.ndarray]]:
        """Yield (index, audio) in submission order, each as soon as it and all earlier sentences are done."""
        n = len(id_rows)
        out: List[Optional[np.ndarray]] = [None] * n
        cv = threading.Condition()

        def on_done(i: int, a: np.ndarray):
            with cv:
                out[i] = a
                cv.notify_all()

        threads, errors = self._run(id_rows, seed, on_done)
        for i in range(n):
            with cv:
                while out[i] is None:
                    if errors:
                        raise errors[0]
                    if not any(t.is_alive() for t in threads) and out[i] is None:
                        raise errors[0] if errors else RuntimeError("scheduler workers exited without a result")
                    cv.wait(timeout=0.05)
                a = out[i]
            yield i, a  # type: ignore[misc]
        for t in threads:
            t.join()
