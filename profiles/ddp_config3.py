"""BASELINE configs[2]: full YOLOv8-L + lowlight_recovery training step (the reference's ASFF yolov8l.yaml), bf16 autocast,
batch 8 per GPU (64 over 8 GPUs), DDP -- with layer 0 swapped for the B200 drop-in, beside the unmodified reference.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        profiles/ddp_config3.py [--steps 10] [--batch 8] [--size 640] [--out gpurun_out/config3.json]

Everything except layer 0 is the UNMODIFIED reference from baseline/_ref (nn/tasks.py DetectionModel, utils/loss.py
RcoveryDetectionLoss, the trainer's own step recipe: autocast -> loss * world_size -> backward; engine/trainer.py:223,
330-340) wrapped in torch DDP exactly as the trainer does (find_unused_parameters=False by default).  Synthetic data
(uint8 batch -> preprocess_batch -> one box per image).  Reports ms/step (CUDA events, max over ranks) for both arms, the
loss of both arms on the first step (same weights, same batch) and whether every parameter received a finite gradient.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    argv, sys.argv = sys.argv, sys.argv[:1]

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from baseline import reference_runtime as R
    ref = R.load_ultralytics()
    import dedark_yolo_b200 as dd
    import dedark_yolo_b200.integrate as it
    from torch.nn.parallel import DistributedDataParallel as DDP

    B, S = args.batch, args.size
    gen = torch.Generator().manual_seed(100 + rank)
    u8 = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=gen)

    def make_batch(preprocess):
        batch = {"img": u8.clone().to(dev)}
        batch = preprocess(batch)
        batch["cls"] = torch.tensor([[float(i % 3)] for i in range(B)], device=dev)
        batch["bboxes"] = torch.tensor([[0.5, 0.5, 0.3, 0.3]] * B, device=dev)
        batch["batch_idx"] = torch.arange(B, dtype=torch.float32, device=dev)
        return batch

    def ref_preprocess(batch):  # models/yolo/detect/train.py:72,103,108 with lowlight_FLAG and no dedark branch
        batch["clean_img"] = batch["img"].float() / 255
        batch["img"] = torch.pow(batch["clean_img"], 15.0)
        batch["recovery_loss_batch"] = torch.nn.functional.mse_loss(batch["img"], batch["clean_img"])
        return batch

    def our_preprocess(batch):
        return dd.preprocess_batch(batch, dev, dark_param=15.0, lowlight_FLAG=True, dedark_FLAG=False)

    def build(state=None):
        torch.manual_seed(0)
        model = ref.DetectionModel(ref.yaml_l, nc=3, verbose=False)
        if state is not None:
            model.load_state_dict(state)
        model.args = ref.get_cfg(ref.DEFAULT_CFG)
        model = model.to(dev).train()
        return model

    def run(model, preprocess, name):
        net = DDP(model, device_ids=[local]) if world > 1 else model
        opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        first_loss, first_items = None, None
        n0 = dd.launch_count()
        for i in range(args.warmup + args.steps):
            if i == args.warmup:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                ev[0].record()
            batch = make_batch(preprocess)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss, items = net(batch)
                if world > 1:
                    loss = loss * world           # engine/trainer.py:334-335
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if first_loss is None:
                first_loss, first_items = float(loss.detach()) / max(world, 1), [float(v) for v in items]
                bad = [k for k, p in model.named_parameters() if p.grad is None or not torch.isfinite(p.grad).all()]
            opt.step()
        ev[1].record()
        torch.cuda.synchronize(dev)
        ms = ev[0].elapsed_time(ev[1]) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return {"arm": name, "ms_per_step": ms, "images_per_s": world * B / (ms * 1e-3), "first_loss": first_loss, "first_items": first_items,
                "params_without_finite_grad": len(bad), "dedark_kernel_launches": dd.launch_count() - n0}

    stock = build()
    state = {k: v.detach().clone() for k, v in stock.state_dict().items()}
    res_ref = run(stock, ref_preprocess, "reference (stock lowlight_recovery, PyTorch eager)")
    del stock
    torch.cuda.empty_cache()
    it.install()
    swapped = build(state)
    assert isinstance(swapped.model[0], dd.lowlight_recovery)
    res_new = run(swapped, our_preprocess, "layer 0 = dedark_yolo_b200.lowlight_recovery, preprocess_batch = dd.preprocess_batch")
    it.uninstall()
    if rank == 0:
        out = {"config": f"configs[2]: YOLOv8-L (ASFF yolov8l.yaml, nc=3) + lowlight_recovery, bf16 autocast, batch {B}/GPU x {world} GPUs = {B * world}, "
                         f"{S}x{S}, DDP, synthetic data", "world": world, "steps": args.steps, "warmup": args.warmup,
               "reference": res_ref, "dropin": res_new,
               "speedup_whole_step": res_ref["ms_per_step"] / res_new["ms_per_step"],
               "loss_rel_diff_first_step": abs(res_new["first_loss"] - res_ref["first_loss"]) / abs(res_ref["first_loss"])}
        line = json.dumps(out)
        print(line)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
