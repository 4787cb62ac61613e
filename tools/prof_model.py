import sys, json, torch
sys.path.insert(0, "/root/repo")
import yolo_sod_b200
from yolo_sod_b200 import synth
from yolo_sod_b200.model import DetectionModel
name = sys.argv[1]; B = int(sys.argv[2])
m = DetectionModel(name, dtype=torch.bfloat16)
x = synth.synth_images(B, 640, seed=0).cuda()
m(x); torch.cuda.synchronize()
prog = m.program(B, 640, 640)
t = prog.profile(iters=2)
tot = sum(v["ms"] for v in t.values())
print(name, B, "sum ms", round(tot, 3))
for k, v in sorted(t.items(), key=lambda kv: -kv[1]["ms"])[:8]:
    print(f"  {k:28s} {v['ms']:.3f} ms n={v['launches']}")
for o in sorted(prog.last_per_op, key=lambda o: -o["ms"])[:14]:
    print(f"  {o['ms']*1e3:8.1f} us {o['kernel'][5:20]:16s} {o['desc'][:110]}")
