"""Timing of one training step (labels -> train-mode forward -> loss -> backward -> Adam) on synthetic pairs.
   python scripts/train_step_time.py [batch n m stages iters]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openglue_b200 import SuperGlue, criterion, generate_gt_matches, _cabi
from openglue_b200.synthetic import default_config, synthetic_pairs, synthetic_state_dict

batch, n, m, stages, iters = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (8, 1024, 1024, 9, 20)))
dev = torch.device('cuda:0')
cfg = default_config(descriptor_dim=256, num_stages=stages, num_iters=iters)
cfg['precision'] = os.environ.get('OG_TRAIN_PREC', 'tf32x3')
model = SuperGlue(cfg)
model.load_state_dict(synthetic_state_dict(cfg, seed=5), strict=True)
model = model.to(dev).train()
pairs = synthetic_pairs(batch, n, m, 256, 1, family='planted', seed=21)
pairs = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in pairs.items()}
H = torch.tensor([[0.9, 0.0, 20.0], [0.0, 0.9, 20.0], [0.0, 0.0, 1.0]], device=dev).repeat(batch, 1, 1)
raw = {'transformation': {'type': ['perspective'] * batch, 'H': H}, 'image0_size': pairs['image0_size'], 'image1_size': pairs['image1_size']}
f0 = {'keypoints': pairs['keypoints0'], 'side_info': pairs['side_info0'], 'local_descriptors': pairs['local_descriptors0']}
f1 = {'keypoints': pairs['keypoints1'], 'side_info': pairs['side_info1'], 'local_descriptors': pairs['local_descriptors1']}
opt = torch.optim.Adam(model.parameters(), lr=1e-4)

def step(parts=None):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    data, y_true = generate_gt_matches(raw, f0, f1, 3.0, 5.0)
    ev[1].record()
    y_pred = model(data)
    loss = criterion(y_true, y_pred, margin=None)['loss']
    ev[2].record()
    opt.zero_grad()
    loss.backward()
    ev[3].record()
    opt.step()
    ev[4].record()
    if parts is not None:
        torch.cuda.synchronize()
        parts.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
parts = []
t0 = time.time()
for _ in range(5):
    l = step(parts)
torch.cuda.synchronize()
wall = (time.time() - t0) / 5 * 1e3
p = [sum(x[i] for x in parts) / len(parts) for i in range(4)]
print('train step: batch %d n %d m %d stages %d iters %d precision %s' % (batch, n, m, stages, iters, cfg['precision']))
print('   labels %.2f ms | forward + loss %.2f ms | backward %.2f ms | Adam %.2f ms | total (device) %.2f ms | wall %.2f ms | %.1f pairs/s | loss %.4f'
      % (p[0], p[1], p[2], p[3], sum(p), wall, batch / (wall * 1e-3), float(l.detach())))

# the same step captured once and replayed as one CUDA graph (openglue_b200.training.GraphedTrainStep)
from openglue_b200.training import GraphedTrainStep
data, y_true = generate_gt_matches(raw, f0, f1, 3.0, 5.0)
g = GraphedTrainStep(model, data, y_true)
for _ in range(2):
    g(data, y_true); opt.step()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
t0 = time.time()
tg = ta = 0.0
for _ in range(5):
    e0.record(); l = g(data, y_true)['loss']; e1.record(); opt.step(); e2.record()
    torch.cuda.synchronize(); tg += e0.elapsed_time(e1); ta += e1.elapsed_time(e2)
wall = (time.time() - t0) / 5 * 1e3
print('   as ONE CUDA graph: forward + loss + backward %.2f ms | Adam %.2f ms | wall %.2f ms | %.1f pairs/s | loss %.4f'
      % (tg / 5, ta / 5, wall, batch / (wall * 1e-3), float(l)))
