"""Both CLIs on 1 GPU and under torchrun on N GPUs: the output files must be identical.
usage (under gpurun --gpus 2): python scripts/multi_gpu_cli_check.py 2"""
import json, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waymo_2d_tracking_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
scene = synth.make_scene(synth.SynthConfig(n_segments=5, n_frames=40, n_submissions=3, objects_per_frame=60.0, seed=77))
env = dict(os.environ, PYTHONPATH=root)
with tempfile.TemporaryDirectory() as tmp:
    files = []
    for k, sub in enumerate(scene.submissions):
        p = os.path.join(tmp, "sub%d.json" % k)
        json.dump(synth.to_json_list(scene, sub), open(p, "w"))
        files.append(p)
    gt = os.path.join(tmp, "images.json")
    json.dump([], open(gt, "w"))
    run = lambda cmd: subprocess.run(cmd, check=True, env=env, cwd=root, capture_output=True, text=True)
    torchrun = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                "--master-addr", "127.0.0.1", "--master-port", "29533"]
    ens_args = ["-m", "soft_nms", "--min-score=0.01", "--soft-nms-cut=0.9"]
    run([sys.executable, "-m", "waymo_2d_tracking_b200.detnet.ensemble"] + files + ["-o", os.path.join(tmp, "ens1.json")] + ens_args)
    run(torchrun + ["-m", "waymo_2d_tracking_b200.detnet.ensemble"] + files + ["-o", os.path.join(tmp, "ensN.json")] + ens_args)
    trk_args = ["--ground-truth", gt, "--input", os.path.join(tmp, "ens1.json"), "--max-age=2", "--min-hits=0"]
    run([sys.executable, "-m", "waymo_2d_tracking_b200.tracking.track"] + trk_args + ["--output", os.path.join(tmp, "trk1.json")])
    run(torchrun + ["-m", "waymo_2d_tracking_b200.tracking.track"] + trk_args + ["--output", os.path.join(tmp, "trkN.json")])
    e1, eN = json.load(open(os.path.join(tmp, "ens1.json"))), json.load(open(os.path.join(tmp, "ensN.json")))
    t1, tN = json.load(open(os.path.join(tmp, "trk1.json"))), json.load(open(os.path.join(tmp, "trkN.json")))
    print("ensemble rows %d, identical on %d GPUs: %s" % (len(e1), n, e1 == eN))
    print("track rows %d, ids up to %s, identical on %d GPUs: %s" % (len(t1), max(int(r['object_id']) for r in t1), n, t1 == tN))
    assert e1 == eN and t1 == tN
