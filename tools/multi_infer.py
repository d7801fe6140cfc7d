"""The reference's `tasks/run.py --infer` seam on N GPUs: fabricates an experiment tree in the
reference's on-disk formats (tests/fake_exp.py), runs `python -m dict_tts_b200.run --exp_name ... --infer` under torchrun
with one rank per GPU, and checks what the ranks left behind: ONE meta.csv with every utterance once, in dataset order,
and one int16 wav per row.  Prints one JSON line.  Run through tools/gpu_multi.sh."""
import argparse
import csv
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import fake_exp  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--items", type=int, default=0)
    a = ap.parse_args()
    n_items = a.items or 4 * a.gpus + 3
    root = tempfile.mkdtemp(prefix="dtts_multi_")
    exp = fake_exp.write(root, n_items=n_items)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr",
           "127.0.0.1", "--master-port", "29533", "-m", "dict_tts_b200.run", "--exp_name", exp["exp"], "--infer", "--hparams",
           "b200_max_sentences=2,gen_dir_name=multi"]
    t0 = time.time()
    out = subprocess.run(cmd, cwd=exp["root"], env=env, capture_output=True, text=True, timeout=900)
    secs = time.time() - t0
    res = dict(gpus=a.gpus, items=n_items, rc=out.returncode, seconds=round(secs, 1))
    if out.returncode == 0:
        gen = os.path.join(exp["work_dir"], "generated_3000_multi")
        with open(os.path.join(gen, "meta.csv")) as f:
            rows = list(csv.DictReader(f))
        names = [r["item_name"] for r in rows]
        wavs = [os.path.exists(os.path.join(gen, "wavs", r["wav_fn_pred"] + ".wav")) for r in rows]
        res.update(rows=len(rows), dataset_order=names == [f"fake_{i:03d}" for i in range(n_items)], wavs=sum(wavs),
                   ok=len(rows) == n_items and all(wavs) and names == [f"fake_{i:03d}" for i in range(n_items)])
    else:
        res.update(ok=False, stderr=out.stderr[-1500:])
    print(json.dumps(res))
    sys.exit(0 if res["ok"] else 1)
