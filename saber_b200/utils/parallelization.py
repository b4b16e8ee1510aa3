"""Tomogram-level data parallelism over the GPUs of one node — the `GPUPool` contract the reference's batch commands are
written against (REF saber/utils/parallelization.py:15-449):

* `GPUPool(init_fn=..., init_args=...)` loads one model set per GPU (`init_fn(gpu_id, *init_args, **init_kwargs)`);
* `execute(func, tasks, task_ids)` sends task i to GPU `i % n_gpus` and calls `func(*args, gpu_id=g, models=models[g],
  **kwargs)`; a task is a kwargs dict, an `(args, kwargs)` pair, an argument list / tuple, or a single argument;
* the result is one dict per task, sorted by task id: `{'success': True, 'task_id', 'gpu_id', 'processing_time', 'result'}`
  or `{'success': False, 'task_id', 'gpu_id', 'error'}` — a failing task never takes the pool down.

Written for this backend rather than transcribed: every GPU gets ONE worker thread that owns that device (tasks of a GPU
run in submission order, so no per-GPU lock is needed and a device is never driven from two threads), which is the
configuration `saber_b200.lib` / `ops` are built for (per-thread current device, thread-local CUDA-graph capture, per-device
function-attribute caches). The reference's second mode (spawned worker processes) corresponds to this package's
one-process-per-GPU launch (`torch.distributed`, `saber_b200/dist.py`, `bench.py --gpus N`) and is not duplicated here.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import Any, Callable, Dict, List, Optional

import torch


def _split_task(task):
    if isinstance(task, dict):
        return (), dict(task)
    if isinstance(task, tuple) and len(task) == 2 and isinstance(task[1], dict):
        return tuple(task[0]) if isinstance(task[0], (list, tuple)) else (task[0],), dict(task[1])
    if isinstance(task, (list, tuple)):
        return tuple(task), {}
    return (task,), {}


class GPUPool:
    def __init__(self, approach: str = "threading", init_fn: Optional[Callable] = None, init_args: tuple = (),
                 init_kwargs: Optional[dict] = None, verbose: bool = True, n_gpus: Optional[int] = None):
        if approach != "threading":
            raise ValueError("saber_b200 GPUPool runs worker threads (approach='threading'); for one process per GPU use "
                             "torch.distributed (saber_b200/dist.py, bench.py --gpus N)")
        if not torch.cuda.is_available():
            raise RuntimeError("saber_b200 GPUPool needs CUDA devices (no CPU fallback)")
        self.approach = approach
        self.n_gpus = torch.cuda.device_count() if n_gpus is None else min(int(n_gpus), torch.cuda.device_count())
        self.init_fn, self.init_args, self.init_kwargs = init_fn, tuple(init_args), dict(init_kwargs or {})
        self.verbose = verbose
        self.models: Dict[int, Any] = {}
        self._loaded = False

    # ---- models: once per GPU, before the first task -----------------------------------------------------------------
    def _load_models(self):
        for g in range(self.n_gpus):
            torch.cuda.set_device(g)
            t0 = time.time()
            self.models[g] = self.init_fn(g, *self.init_args, **self.init_kwargs) if self.init_fn else None
            if self.verbose and self.init_fn:
                print(f"GPU {g}: models loaded in {time.time() - t0:.1f}s, "
                      f"{torch.cuda.memory_allocated(g) / 1e9:.1f} GB allocated")
        self._loaded = True

    def start(self):
        """Kept for interface parity (the reference starts worker processes here); models load on first use."""
        if not self._loaded:
            self._load_models()

    # ---- execution -------------------------------------------------------------------------------------------------------
    def execute(self, func: Callable, tasks: List[Any], task_ids: Optional[List] = None,
                progress_desc: str = "Processing") -> List[Dict]:
        if not tasks:
            return []
        if task_ids is None:
            task_ids = list(range(len(tasks)))
        self.start()
        lanes = [queue.SimpleQueue() for _ in range(self.n_gpus)]
        for i, (tid, task) in enumerate(zip(task_ids, tasks)):
            lanes[i % self.n_gpus].put((tid, *_split_task(task)))
        results: List[Dict] = []
        lock = threading.Lock()

        def drain(g: int):
            torch.cuda.set_device(g)
            while True:
                try:
                    tid, args, kwargs = lanes[g].get_nowait()
                except queue.Empty:
                    return
                kwargs["gpu_id"] = g
                if self.models.get(g) is not None:
                    kwargs["models"] = self.models[g]
                t0 = time.time()
                try:
                    rec = {"success": True, "task_id": tid, "gpu_id": g, "result": func(*args, **kwargs),
                           "processing_time": time.time() - t0}
                except Exception as e:  # a failing task is reported, not raised (REF :130-136)
                    rec = {"success": False, "task_id": tid, "gpu_id": g, "error": str(e)}
                with lock:
                    results.append(rec)
                    if self.verbose:
                        state = f"{rec['processing_time']:.1f}s" if rec["success"] else f"FAILED: {rec['error']}"
                        print(f"{progress_desc}: {len(results)}/{len(tasks)}  task {tid} on GPU {g}  {state}")

        threads = [threading.Thread(target=drain, args=(g,), name=f"saber_b200_gpu{g}") for g in range(self.n_gpus)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if self.verbose:
            self._print_stats(results)
        try:
            return sorted(results, key=lambda r: r.get("task_id", 0))
        except TypeError:  # task ids that do not order (mixed types): submission order
            pos = {id(t): i for i, t in enumerate(task_ids)}
            return sorted(results, key=lambda r: pos.get(id(r.get("task_id")), 0))

    def _print_stats(self, results):
        ok = [r for r in results if r["success"]]
        bad = [r for r in results if not r["success"]]
        print(f"GPUPool: {len(results)} tasks, {len(ok)} succeeded, {len(bad)} failed")
        for r in bad:
            print(f"  - {r['task_id']}: {r['error']}")
        for g in range(self.n_gpus):
            mine = [r["processing_time"] for r in ok if r["gpu_id"] == g]
            if mine:
                print(f"  GPU {g}: {len(mine)} tasks, avg {sum(mine) / len(mine):.2f}s/task")

    def shutdown(self):
        """Nothing outlives `execute` (worker threads are joined there); the models are released."""
        self.models.clear()
        self._loaded = False

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.shutdown()


def gpu_map(func: Callable, tasks: List[Any], approach: str = "threading", n_gpus: Optional[int] = None,
            init_fn: Optional[Callable] = None, init_args: tuple = (), init_kwargs: Optional[dict] = None,
            verbose: bool = True) -> List[Dict]:
    """REF :451-470: one-shot pool."""
    with GPUPool(approach=approach, init_fn=init_fn, init_args=init_args, init_kwargs=init_kwargs, verbose=verbose,
                 n_gpus=n_gpus) as pool:
        return pool.execute(func, tasks)
