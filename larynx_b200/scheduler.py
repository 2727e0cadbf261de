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

import queue
import threading
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import dispatch

SynthFn = Callable[[List[np.ndarray], Optional[int]], List[np.ndarray]]


class BatchScheduler:
    """Workers live as long as the scheduler: each owns ONE call context (CUDA stream + workspace arenas) that is created
    on its first micro-batch and reused by every later request, so serving pays no allocation or device-wide sync per
    request (a thread-per-request design would build and free gigabytes of workspaces each time)."""

    def __init__(self, tts_model=None, vocoder_model=None, glue=None, max_batch: int = 16, max_cost: Optional[int] = 4096,
                 workers: int = 2, noise_scale: float = 0.667, length_scale: float = 1.0, denoiser_strength: float = 0.0,
                 synth: Optional[SynthFn] = None):
        """`synth(rows, seed) -> list of int16 arrays` replaces the device path (tests); by default it is
        `host.synthesize_batch` on the given models.  `seed` None = a fresh noise stream per micro-batch."""
        assert max_batch >= 1 and workers >= 1
        self.max_batch, self.max_cost, self.workers = int(max_batch), max_cost, int(workers)
        if synth is None:
            assert tts_model is not None and vocoder_model is not None and glue is not None
            from . import engine, host

            device = tts_model.model.device
            tls = threading.local()

            def synth(rows: List[np.ndarray], seed: Optional[int]) -> List[np.ndarray]:
                call = getattr(tls, "call", None)
                if call is None:  # once per worker thread, kept for the scheduler's lifetime
                    call = tls.call = engine.Call(device)
                return host.synthesize_batch(tts_model, vocoder_model, glue, rows, noise_scale=noise_scale,
                                             length_scale=length_scale, seed=seed, denoiser_strength=denoiser_strength,
                                             call=call)["audio"]

        self._synth = synth
        self._jobs: "queue.Queue" = queue.Queue()
        self._threads = [threading.Thread(target=self._worker, daemon=True, name=f"b200tts-sched-{i}") for i in range(self.workers)]
        self._closed = False
        for t in self._threads:
            t.start()

    def close(self):
        """Stop the workers (their call contexts are released with them)."""
        if not self._closed:
            self._closed = True
            for _ in self._threads:
                self._jobs.put(None)
            for t in self._threads:
                t.join()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _worker(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            if callable(job):
                job()
                continue
            req, k = job
            req.run(k, self._synth)

    def reserve(self, id_rows: Sequence[np.ndarray]):
        """Size every worker's workspaces for a request like `id_rows` before it matters: each worker runs the costliest
        micro-batch of the plan once (results dropped).  Which worker picks which micro-batch is not deterministic, so
        without this the first long micro-batch a worker meets grows its arenas (cudaFree + cudaMalloc of gigabytes: tens of
        milliseconds and a device-wide sync that also stalls the other workers) — seen as a 2x swing of the batch-32
        throughput from run to run.  A serving process calls it once at start-up with its longest expected request."""
        assert not self._closed, "scheduler is closed"
        rows = [np.asarray(r, dtype=np.int64).reshape(-1) for r in id_rows]
        plan = self.plan(rows)
        if not plan:
            return
        big = max(plan, key=lambda idx: len(idx) * max(len(rows[i]) for i in idx))
        batch = [rows[i] for i in big]
        gate = threading.Barrier(self.workers + 1)
        errors: List[BaseException] = []

        def job():
            try:
                self._synth(batch, 0)
            except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
                errors.append(e)
            finally:
                gate.wait()  # nobody takes a second job before every worker has had its one

        for _ in range(self.workers):
            self._jobs.put(job)
        gate.wait()
        if errors:
            raise errors[0]

    # ------------------------------------------------------------------------------------------------------------
    def plan(self, id_rows: Sequence[np.ndarray]) -> List[List[int]]:
        """Micro-batches (lists of sentence indices), longest first; empty sentences are left out."""
        lens = [len(r) for r in id_rows]
        keep = [i for i, n in enumerate(lens) if n > 0]
        buckets = dispatch.bucket_by_length([lens[i] for i in keep], self.max_batch, self.max_cost)
        return [[keep[j] for j in b] for b in buckets][::-1]

    def _submit(self, id_rows: Sequence[np.ndarray], seed: Optional[int], on_done: Callable[[int, np.ndarray], None]) -> "_Request":
        assert not self._closed, "scheduler is closed"
        rows = [np.asarray(r, dtype=np.int64).reshape(-1) for r in id_rows]
        for i, r in enumerate(rows):
            if len(r) == 0:
                on_done(i, np.zeros(0, dtype=np.int16))
        req = _Request(rows, self.plan(rows), seed, on_done)
        for k in range(len(req.plan)):
            self._jobs.put((req, k))
        return req

    def synthesize(self, id_rows: Sequence[np.ndarray], seed: Optional[int] = None) -> List[np.ndarray]:
        """int16 audio of every sentence, in submission order.  `seed` fixes the noise (micro-batch k uses seed + k)."""
        out: List[Optional[np.ndarray]] = [None] * len(id_rows)

        def on_done(i: int, a: np.ndarray):
            out[i] = a

        req = self._submit(id_rows, seed, on_done)
        req.wait()
        if req.errors:
            raise req.errors[0]
        assert all(a is not None for a in out)
        return out  # type: ignore[return-value]

    def stream(self, id_rows: Sequence[np.ndarray], seed: Optional[int] = None) -> Iterator[Tuple[int, np.ndarray]]:
        """Yield (index, audio) in submission order, each as soon as it and all earlier sentences are done."""
        n = len(id_rows)
        out: List[Optional[np.ndarray]] = [None] * n
        cv = threading.Condition()

        def on_done(i: int, a: np.ndarray):
            with cv:
                out[i] = a
                cv.notify_all()

        req = self._submit(id_rows, seed, on_done)
        for i in range(n):
            with cv:
                while out[i] is None:
                    if req.errors:
                        raise req.errors[0]
                    if req.finished() and out[i] is None:
                        raise req.errors[0] if req.errors else RuntimeError("scheduler finished without a result")
                    cv.wait(timeout=0.05)
                a = out[i]
            yield i, a  # type: ignore[misc]
        req.wait()


class _Request:
    """One synthesize()/stream() call: its micro-batch plan and completion state."""

    def __init__(self, rows, plan, seed, on_done):
        self.rows, self.plan, self.seed, self.on_done = rows, plan, seed, on_done
        self.errors: List[BaseException] = []
        self._left = len(plan)
        self._cv = threading.Condition()

    def run(self, k: int, synth: SynthFn):
        try:
            if not self.errors:
                idx = self.plan[k]
                audio = synth([self.rows[i] for i in idx], None if self.seed is None else self.seed + k)
                assert len(audio) == len(idx)
                for i, a in zip(idx, audio):
                    self.on_done(i, a)
        except BaseException as e:  # noqa: BLE001 - re-raised on the caller's thread
            self.errors.append(e)
        finally:
            with self._cv:
                self._left -= 1
                self._cv.notify_all()

    def finished(self) -> bool:
        with self._cv:
            return self._left <= 0

    def wait(self):
        with self._cv:
            while self._left > 0:
                self._cv.wait(timeout=0.05)
