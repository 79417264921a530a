import sys, os, torch
sys.path.insert(0, ".")
from adaprompt_b200.ldm_lite import SD15_UNET_CONFIG, LatentDiffusionLite
from adaprompt_b200.unet import UNetModel
from adaprompt_b200.weights import spec_of, synth_state_dict
from adaprompt_b200.ddim import DDIMSampler
from oracle.golden_inputs import ddim_inputs
with torch.device("meta"):
    m = UNetModel(**SD15_UNET_CONFIG)
m = m.to_empty(device="cuda"); m.load_state_dict(synth_state_dict(spec_of(m), 1234)); m.eval()
gold = torch.load("tests/golden/ddim_traj.pt")["s10_32_g4_1"]
S, shape, cond, uncond, gs, x_T = ddim_inputs("s10_32_g4_1")
cond = (cond[0].cuda(), cond[1], cond[2]); uncond = (uncond[0].cuda(), uncond[1], uncond[2])
def rel(a, b): a, b = a.float().cpu(), b.float().cpu(); return ((a-b).norm()/b.norm()).item()
import sys as _s
name = _s.argv[1] if len(_s.argv) > 1 else "s10_32_g4_1"
gold = torch.load("tests/golden/ddim_traj.pt")[name]
S, shape, cond, uncond, gs, x_T = ddim_inputs(name)
cond = (cond[0].cuda(), cond[1], cond[2]); uncond = (uncond[0].cuda(), uncond[1], uncond[2])
LOG = max(1, S // 10)
for graph in (False, True, True):
    model = LatentDiffusionLite(m).cuda()
    sm = DDIMSampler(model, use_cuda_graph=graph)
    s, inter = sm.sample(S, 1, list(shape[1:]), conditioning=cond, unconditional_conditioning=uncond, guidance_scale=gs, eta=0.0, x_T=x_T.cuda(), verbose=False, log_every_t=LOG)
    torch.cuda.synchronize()
    print("graph", graph, "final", rel(s, gold["samples"]), "steps", ["%.1e" % rel(a, b) for a, b in zip(inter["x_inter"][1:], gold["x_inter"][1:])], flush=True)
    print("   pred", ["%.1e" % rel(a, b) for a, b in zip(inter["pred_x0"][1:], gold["pred_x0"][1:])], flush=True)
