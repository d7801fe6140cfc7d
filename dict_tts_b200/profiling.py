"""Stage ranges under the reference's own Timer names (tracing row).

The reference brackets its inference stages with ``utils.Timer(name, enable=hparams['profile_infer'])``
(utils/__init__.py:260-281): 'encoder' and, inside it, 'dict_encoder' (modules/dict_tts/model.py:50,86), 'fvae'
(model.py:57) and 'hifigan' (vocoders/hifigan.py:59).  ``Timer`` here has the same constructor and the same
accumulate-and-print behaviour when enabled (synchronise, wall clock, ``[Timer] name: total``); enabled or not it also
opens an NVTX range of that name, so a profiler timeline of this engine reads like one of the reference.
"""
import time

import torch


class Timer:
    timer_map = {}

    def __init__(self, name: str, enable: bool = False):
        Timer.timer_map.setdefault(name, 0.0)
        self.name, self.enable = name, enable
        self._nvtx = False

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
            self._nvtx = True
            if self.enable:
                torch.cuda.synchronize()
        if self.enable:
            self.t = time.time()
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        if self.enable:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            Timer.timer_map[self.name] += time.time() - self.t
            print(f"[Timer] {self.name}: {Timer.timer_map[self.name]}")
        if self._nvtx:
            torch.cuda.nvtx.range_pop()
        return False
