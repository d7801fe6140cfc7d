"""``python -m dict_tts_b200.run --config <yaml> --exp_name <name> --infer [--hparams k=v,...]``

Same flags as the reference's ``tasks/run.py`` (:35-42 with utils/hparams.py:27-38).  The task class comes from
``hparams['task_cls']``; the reference's own class names map onto the B200 task so an unmodified config runs here.
Multi-GPU: launch under torchrun (one process per GPU); the batches are dealt round-robin across ranks.
"""
import importlib

from . import hparams as hp_mod

_ALIASES = {"tasks.tts.dict_tts.DictTTSTask": "dict_tts_b200.task.B200DictTTSTask",
            "tasks.tts.ps_flow.PortaSpeechFlowTask": "dict_tts_b200.task.B200PortaSpeechTask",
            "tasks.tts.ps_adv.PortaSpeechAdvTask": "dict_tts_b200.task.B200PortaSpeechTask"}


def run_task():
    name = hp_mod.hparams.get("task_cls", "") or "dict_tts_b200.task.B200DictTTSTask"
    name = _ALIASES.get(name, name)
    pkg, cls = name.rsplit(".", 1)
    return getattr(importlib.import_module(pkg), cls).start()


def main(argv=None):
    args = hp_mod.parse_args(argv)
    hp_mod.set_hparams(args.config, args.exp_name, args.hparams, print_hparams=False, infer=args.infer,
                       reset=args.reset)
    if not args.infer:
        raise SystemExit("only --infer is implemented (training is out of scope, DESIGN.md)")
    return run_task()


if __name__ == "__main__":
    main()
