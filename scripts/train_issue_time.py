"""How much of a Stage-1 optimizer step is host issue time?  Times, per micro-batch, the host time until all launches of
micro_backward are issued (no sync) and the device time until they complete."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adaprompt_b200.synthetic import stage1_batch, stage1_stack
from adaprompt_b200.train_cond import Stage1Trainer
dev = torch.device("cuda", 0)
step, params = stage1_stack(dev)
trainer = Stage1Trainer(step, params, world_size=1, accum=2, use_graph="step")
g = torch.Generator().manual_seed(100)
for _ in range(3):
    trainer.optimizer_step([stage1_batch(dev, 4, 64, g) for _ in range(2)])
torch.cuda.synchronize()
for it in range(3):
    batches = [stage1_batch(dev, 4, 64, g) for _ in range(2)]
    torch.cuda.synchronize()
    trainer.bucket.begin_step()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    marks = []
    for b in batches:
        ta = time.perf_counter()
        c = step.context(b["face_embs"], b["tokens"])
        tb = time.perf_counter()
        from adaprompt_b200.train import GraphedUNetLoss
        gl = step.__dict__["_graphed"]
        loss, grad_c = gl(step.q_sample(b["x0"], b["t"], b["noise"]), b["t"], c.detach(), b["teacher_eps"])
        tc = time.perf_counter()
        c.backward(grad_c * 0.5)
        td = time.perf_counter()
        marks.append((tb - ta, tc - tb, td - tc))
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {it}: host issue {1e3 * (t1 - t0):.1f} ms, device {e0.elapsed_time(e1):.1f} ms, wall to sync {1e3 * (t2 - t0):.1f} ms")
    for m in marks:
        print("   context fwd issue %.1f ms | graph issue %.1f ms | context bwd issue %.1f ms" % tuple(1e3 * v for v in m))
    trainer.bucket.allreduce(1)
    gn = trainer.bucket.clip_(0.5)
    t3 = time.perf_counter()
    trainer.optimizer.step(trainer.bucket.flat)
    torch.cuda.synchronize()
    print(f"   clip + prodigy {1e3 * (time.perf_counter() - t3):.1f} ms")
# device time of the pieces, each synchronised
b = stage1_batch(dev, 4, 64, g)
def dev_ms(f):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = f(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1), r
ms_c, c = dev_ms(lambda: step.context(b["face_embs"], b["tokens"]))
gl = step.__dict__["_graphed"]
ms_g, (loss, grad_c) = dev_ms(lambda: gl(step.q_sample(b["x0"], b["t"], b["noise"]), b["t"], c.detach(), b["teacher_eps"]))
ms_b, _ = dev_ms(lambda: c.backward(grad_c * 0.5))
print(f"synchronised pieces (wall incl. host): context fwd {ms_c:.1f} ms, UNet graph {ms_g:.1f} ms, context bwd {ms_b:.1f} ms")
