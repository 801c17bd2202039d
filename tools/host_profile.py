"""cProfile of the host side of the train step (where does the enqueue time go?).
Usage: python tools/host_profile.py [steps]"""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import s4former_b200 as s4
from s4former_b200 import configs, ops
from s4former_b200.runner import TrainStep
from s4former_b200.utils.synthetic import make_batch, fresh_metas
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device('cuda', 0)
ops.set_compute_dtype(torch.bfloat16)
cfg = configs.setr_pup_deit_base('ours', 512, 21, norm='SyncBN')
torch.manual_seed(1999)
model = s4.build_segmentor(cfg)
model.init_weights()
model.backbone_ema.load_state_dict(model.backbone.state_dict())
model.decode_head_ema.load_state_dict(model.decode_head.state_dict())
model = model.to(dev).train()
step = TrainStep(model)
img, gt, metas = make_batch(8, 8, 512, 21, seed=1999)
img, gt = img.to(dev), gt.to(dev)
for i in range(3):
    step(img, fresh_metas(metas), gt, i, sync=False)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(steps):
    step(img, fresh_metas(metas), gt, 3 + i, sync=False)
pr.disable()
torch.cuda.synchronize()
for key in ('tottime', 'cumtime'):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
