"""Headline benchmark: DeepLabv3+ (Xception, ds=16) bf16 training images/sec at 512x512.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one full training step (forward, focal+dice loss, backward, gradient all-reduce
over NCCL for N>1, fused Adam) on a synthetic batch of B images per GPU (weak scaling).  Rank 0
prints ONE JSON line (contract in the round brief).  ``--impl reference`` times the UNMODIFIED reference's own
``fit_one_epoch`` (installed under baseline/_ref by oracle/build_ref.py) on the host cores instead; when that install
is absent it falls back to the oracle port and says so (``kind``).  The product arm never imports ``oracle/``: the
reference legs (``cpu_baseline``, ``torch_gpu_baseline``) run ``oracle/ref_runner.py`` in a subprocess.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "deeplabv3plus_xception_512x512_train_images_per_sec"
UNIT = "images/s"
TRAIN_GFLOP_PER_IMAGE = 496.4   # fwd+dgrad+wgrad conv FLOPs, Xception ds=16
CLS_WEIGHTS = [1, 1, 5, 3, 4]   # train.py:274


def synthetic_batch(batch: int, size: int = 512, num_classes: int = 5, seed: int = 0, ignore_frac: float = 0.01):
    """Synthetic inputs per SURVEY.md section 8d (dataloader contract, row L): images U[0,1) fp32 NCHW, masks int64 in
    [0, num_classes) with ``ignore_frac`` of the pixels set to num_classes (ignore_index), one-hot labels [B,H,W,C+1]."""
    import torch
    g = torch.Generator().manual_seed(1000 + seed)
    imgs = torch.rand(batch, 3, size, size, generator=g)
    pngs = torch.randint(0, num_classes, (batch, size, size), generator=g)
    ign = torch.rand(batch, size, size, generator=g) < ignore_frac
    pngs = torch.where(ign, torch.full_like(pngs, num_classes), pngs)
    return imgs, pngs, torch.eye(num_classes + 1)[pngs]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p.get("hbm_gbs", 6650.0), tf_burst=p.get("bf16_tflops", 1590.0),
                    tf_sustained=p.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons during the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def cpu_train_throughput(steps: int, warmup: int, budget_s: float, size: int = 512, batch: int = 2):
    """The reference's train step (fwd + focal + dice + backward + Adam, utils_fit.py:60-90) run
    by the CPU oracle port with all host threads.  Returns (images/s, steps actually timed, cores)."""
    import torch
    from oracle import deeplab_ref as O
    from oracle import losses_ref as L
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = O.make_state("xception", 5, 16, seed=0, randomize_bn_stats=False)
    params = {k: v.clone().requires_grad_(True) for k, v in state.items()
              if v.dtype.is_floating_point and "running" not in k}
    bufs = {k: v.clone() for k, v in state.items() if k not in params}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4, betas=(0.9, 0.999))
    cls_w = torch.tensor(CLS_WEIGHTS, dtype=torch.float32)
    imgs, pngs, labels = O.synthetic_batch(batch, size, seed=0)

    def one():
        opt.zero_grad()
        st = dict(bufs); st.update(params)
        y = O.deeplab_forward(imgs, st, "xception", 16, True, dropout=True)
        loss = L.focal_loss(y, pngs, cls_w, 5) + L.dice_loss(y, labels)
        with torch.no_grad():
            L.f_score(y, labels)
        loss.backward()
        opt.step()
        return float(loss)

    t_start = time.perf_counter()
    for _ in range(warmup):
        one()
        if time.perf_counter() - t_start > budget_s / 3:
            break
    times = []
    for _ in range(steps):
        t0 = time.perf_counter(); one(); times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    sec = sum(times) / len(times)
    return batch / sec, len(times), cores, sec


def cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ref_subprocess(argv, timeout_s: float):
    """Run oracle/ref_runner.py (the unmodified reference's fit_one_epoch on synthetic batches) in a child process and
    return its JSON line, or {"error": ...}.  A child keeps the reference's top-level ``nets`` / ``utils`` packages and its
    cuDNN state out of this process."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py")] + [str(a) for a in argv]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"error": "timed out after %.0f s" % timeout_s}
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                break
    return {"error": "rc=%d: %s" % (out.returncode, (out.stderr or out.stdout).strip().splitlines()[-1:] or "no output")}


def reference_installed() -> bool:
    return os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json"))


def cpu_reference_baseline(steps: int, warmup: int, budget_s: float):
    """The reference's CPU path for this workload on the box's host cores: its own fit_one_epoch (fp32 eager, all
    threads) on batch-2 steps at 512x512 (BatchNorm needs >= 2 images; a batch-32 step is ~20 s of CPU work).  Falls back
    to the oracle port when baseline/_ref is not installed."""
    cores = os.cpu_count() or 1
    if reference_installed():
        r = ref_subprocess(["--device", "cpu", "--batch", 2, "--steps", steps, "--warmup", warmup, "--threads", cores,
                            "--budget-s", budget_s], timeout_s=budget_s * 3 + 120)
        if "error" not in r:
            return {"value": r["images_per_s"], "unit": UNIT, "cores": cores, "kind": "reference", "cpu": cpu_model_name(),
                    "sample": "%d timed steps of the unmodified reference's fit_one_epoch (utils_fit.py:31-198; fp32, Adam, "
                              "focal+dice), Xception ds=16, batch 2 at 512x512 (%.2f s/step)" % (r["steps_timed"], r["ms_per_step"] / 1e3),
                    "ms_per_step": r["ms_per_step"], "steps": r["steps_timed"]}
        err = r["error"]
    else:
        err = "baseline/_ref not installed"
    ips, nsteps, cores, sec = cpu_train_throughput(steps, warmup, budget_s=budget_s)
    return {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
            "sample": "%d timed steps of the oracle port's train step, batch 2 at 512x512 (%.2f s/step); unmodified reference "
                      "unavailable: %s" % (nsteps, sec, err), "ms_per_step": sec * 1e3, "steps": nsteps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_baseline(args.steps, min(args.warmup, 1), budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": cb["steps"],
        "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DeepLabv3+ Xception ds=16 training step (fwd + focal+dice + bwd + Adam) at 512x512, 5 classes, "
                               "on the host CPU cores (bounded sample of BASELINE configs[2]: batch 2 per step)",
                   "batch_per_step": 2},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "cpu", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def torch_gpu_baseline(batch: int, size: int):
    """SURVEY section 8d "the real bar": the unmodified reference's fit_one_epoch on THIS B200 under stock PyTorch / cuDNN
    with fp16 autocast + GradScaler (train.py:82, utils_fit.py:92-121), same batch; and the same with
    ``channels_last`` weights and inputs (a tuning the reference does not apply, listed for completeness)."""
    if not reference_installed():
        return {"unavailable": "baseline/_ref not installed on this box"}
    out = {"what": "unmodified reference fit_one_epoch on this GPU, stock PyTorch/cuDNN (cudnn.benchmark as train.py:383), "
                   "per-step time includes its own H2D copies and .item() syncs", "batch": batch, "size": size}
    for key, extra in (("fp16_autocast", []), ("fp16_autocast_channels_last", ["--channels-last"])):
        r = ref_subprocess(["--device", "cuda", "--batch", batch, "--size", size, "--steps", 6, "--warmup", 3, "--fp16"] + extra,
                           timeout_s=420)
        out[key] = r if "error" in r else {"images_per_s": r["images_per_s"], "ms_per_step": r["ms_per_step"],
                                            "steps": r["steps_timed"]}
    return out


# ----------------------------------------------------------------------------- four-modal classifier leg
def classifier_throughput(steps: int, warmup: int, patients: int = 16, size: int = 512, world: int = 1, rank: int = 0,
                          variants=(("imgN", "imgA", "imgL", "cli"),), use_graph: bool = True):
    """BASELINE configs[1] (N=1, 16 patients) / configs[4] (N>1, `patients` per GPU, fusion-head gradients all-reduced
    over NCCL): one training step of the severity classifier on `patients` patients per rank - the colposcopic
    images of each patient (resize 1024 -> 16 patches of 256 -> frozen ResNet-101 encoder, as
    Graph_Structure(data_augmentation).py:136-200 does; resize + split + normalise is one kernel) + the clinical
    node features, then the fusion head's forward, objective (my_train(full).py:309-347), backward, gradient
    all-reduce and Adam on the flat parameters (engine.FusionTrainer).  Inputs are resident in HBM.  `variants` lists
    the modality sets to time (the reference's 2-/3-/4-modal scripts); the first is the headline.  Times are the max
    over ranks; returns patients/s over all ranks with the split encoder / head."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from cervix_b200.engine import FusionTrainer
    from cervix_b200.multimodal.cli_features import AgeNodeFeatures
    from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2, get_edge_index_full, get_edge_index_image
    from cervix_b200.multimodal.patch_encoder import ResNet101Encoder
    torch.manual_seed(0)
    enc = ResNet101Encoder().cuda().eval()
    for m in enc.modules():   # ImageNet weights are not available offline: random init + non-trivial running statistics
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    torch.manual_seed(100 + rank)
    imgs = torch.rand(3, patients, 3, size, size, device="cuda")       # [modality][patient] images of this rank's shard
    ages = torch.randint(20, 81, (patients,))                          # random age scalars (BASELINE north_star inputs)
    cli = AgeNodeFeatures(max_age=100).cuda()(ages, 20, 80)            # [patients, 4, 1024] clinical node rows
    labels = torch.randint(0, 4, (patients,), device="cuda")
    all_edges = {"imgN": get_edge_index_image(), "imgA": get_edge_index_image(), "imgL": get_edge_index_image(),
                 "cli": get_edge_index_full(4)}
    img_index = {"imgN": 0, "imgA": 1, "imgL": 2}
    chunk = 16                                                         # patients per encoder pass (256 patches)
    results = []
    for types in variants:
        types = list(types)
        torch.manual_seed(0)                                           # identical initial head on every rank
        head = fusion_model_mae_2(1024, 512, 512, 0.3, len(types)).cuda().train()
        trainer = FusionTrainer(head, types, lr=1e-4, weight_decay=5e-4 if len(types) == 4 else 1e-3, world_size=world)
        edges = {m: all_edges[m] for m in types}
        rng = np.random.RandomState(rank)
        masks = np.ones((patients, len(types)), dtype=bool)
        masks[np.arange(patients), rng.randint(0, len(types), patients)] = False   # one visible modality per patient
        n_img = sum(1 for m in types if m in img_index)

        def encode():
            feats = {}
            with torch.no_grad():
                for m in types:
                    if m in img_index:
                        f = [enc.encode_images(imgs[img_index[m], i:i + chunk]) for i in range(0, patients, chunk)]
                        feats[m] = torch.cat(f).view(patients, 16, 1024)
            if "cli" in types:
                feats["cli"] = cli
            return feats

        def draw_masks():
            mk = np.ones((patients, len(types)), dtype=bool)
            mk[np.arange(patients), rng.randint(0, len(types), patients)] = False
            return mk

        # the head's step is ~600 launches of a few microseconds: captured once as a CUDA graph (new masks, dropout
        # streams and Adam step count reach every replay through device memory), eager with --no-graph
        for _ in range(max(warmup, 4) if results else 30):   # untimed encoder passes (allocator, lazy weight folding); the
            encode()                                         # first variant also burns in ~1 s so that the clocks / power
                                                             # state left by the segmentation leg do not land in its timing
        torch.cuda.synchronize()
        if use_graph:
            trainer.capture(encode(), edges, labels, masks, warmup=max(warmup, 1))
            head_step = lambda f: trainer.step_graphed(f, labels, draw_masks())      # noqa: E731
        else:
            for _ in range(warmup):
                trainer.step(encode(), edges, labels, masks)
            head_step = lambda f: trainer.step(f, edges, labels, draw_masks())       # noqa: E731
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        enc_ms = head_ms = 0.0
        for _ in range(steps):
            ev[0].record(); f = encode(); ev[1].record(); loss = head_step(f); ev[2].record()
            torch.cuda.synchronize()
            enc_ms += ev[0].elapsed_time(ev[1]); head_ms += ev[1].elapsed_time(ev[2])
        # ---- end to end: the patients' images start in pinned HOST memory every step (fp32 [modality, patient, 3, H, W],
        # what the reference's PIL -> ToTensor path produces); the copy of step i+1 runs on a side stream under the compute
        # of step i (two device buffers); the loss is read back every step
        n_img_all = len([m for m in types if m in img_index])
        e2e_ms = None
        if n_img_all:
            used = sorted(img_index[m] for m in types if m in img_index)
            host = imgs[used].cpu().pin_memory()
            dev_buf = [torch.empty_like(imgs[used]) for _ in range(2)]
            copy_stream = torch.cuda.Stream()
            done = [torch.cuda.Event(), torch.cuda.Event()]
            ready = [torch.cuda.Event(), torch.cuda.Event()]

            def fetch(k):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done[k & 1])
                    dev_buf[k & 1].copy_(host, non_blocking=True)
                    ready[k & 1].record(copy_stream)

            def encode_from(buf):
                feats = {}
                with torch.no_grad():
                    for j, m in enumerate(mm for mm in types if mm in img_index):
                        f = [enc.encode_images(buf[j, i:i + chunk]) for i in range(0, patients, chunk)]
                        feats[m] = torch.cat(f).view(patients, 16, 1024)
                if "cli" in types:
                    feats["cli"] = cli
                return feats

            torch.cuda.synchronize()
            for e in done:
                e.record()
            t0 = time.perf_counter()
            ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ee0.record()
            fetch(0)
            prev = None
            for k in range(steps):
                if k + 1 < steps:
                    fetch(k + 1)
                torch.cuda.current_stream().wait_event(ready[k & 1])
                l = head_step(encode_from(dev_buf[k & 1]))
                done[k & 1].record()
                l = l.clone()
                if prev is not None:
                    float(prev)
                prev = l
            float(prev)
            ee1.record()
            torch.cuda.synchronize()
            e2e_ms = max(ee0.elapsed_time(ee1), (time.perf_counter() - t0) * 1e3) / steps
            h2d = host.numel() * 4
            del host, dev_buf
        t = torch.tensor([enc_ms / steps, head_ms / steps, e2e_ms or 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        enc_ms, head_ms, e2e_ms = float(t[0]), float(t[1]), float(t[2])
        ms = enc_ms + head_ms
        enc_tf = (n_img * patients * 16 * 20.38e9 / (enc_ms * 1e-3) / 1e12) if n_img else None
        res = {"modalities": types, "patients_per_s": world * patients / (ms * 1e-3),
               "images_per_s": world * n_img * patients / (ms * 1e-3), "ms_per_step": ms, "encoder_ms": enc_ms,
               "head_ms": head_ms, "steps": steps, "encoder_tflops_per_gpu": enc_tf, "loss_last_step": float(loss)}
        if e2e_ms:
            res["e2e"] = {"value": world * patients / (e2e_ms * 1e-3), "unit": "patients/s", "ms_per_step": e2e_ms,
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
        if enc_tf:
            res["roofline"] = {"bound": "tensor", "scope": "ResNet-101 patch encoder forward (20.38 GF per 256x256 patch)",
                               "achieved": enc_tf, "peak": load_peaks()["tf_sustained"], "unit": "TFLOP/s",
                               "frac": enc_tf / load_peaks()["tf_sustained"]}
        results.append(res)
        del trainer, head
    out = dict(results[0])
    out["workload"] = ("%d-modal severity classifier train step, %d patients per GPU x %d images %dx%d (ResNet-101 patch "
                       "encoder bf16, frozen) + clinical nodes -> fusion head fwd/bwd + %sAdam (BASELINE configs[%d])"
                       % (len(results[0]["modalities"]), patients, 3, size, size, "NCCL all-reduce + " if world > 1 else "",
                          1 if world == 1 else 4))
    out["global_patients"] = world * patients
    out["head_cuda_graph"] = bool(use_graph)
    if len(results) > 1:
        out["variants"] = results[1:]
    return out


def classifier_cpu_baseline(size: int = 512, reps: int = 1):
    """CPU baseline of the classifier's train step on a BOUNDED sample (one patient = 3 images = 48 patches): the
    reference's own feature extractor - torchvision ``resnet101`` with ``fc = Linear(2048, 1024)`` in eval mode, one
    batch of the 48 ImageNet-normalised 256x256 crops of the 1024x1024 bilinear resize
    (Graph_Structure(data_augmentation).py:136-200) - followed by the oracle port of the fusion head's forward, objective
    and backward for that patient (oracle/fusion_ref.py), all host threads, fp32."""
    import torch
    import torch.nn.functional as F
    import torchvision
    from cervix_b200.multimodal.my_mae_model import fusion_model_mae_2
    from oracle import fusion_ref as FR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    enc = torchvision.models.resnet101(weights=None)
    enc.fc = torch.nn.Linear(2048, 1024)
    enc.eval()
    state = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in
             fusion_model_mae_2(1024, 512, 512, 0.3, 4).state_dict().items()}
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    imgs = torch.rand(3, 3, size, size)
    mask = [False, True, True, True]
    times = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        with torch.no_grad():
            big = F.interpolate(imgs, size=(1024, 1024), mode="bilinear", align_corners=False)
            patches = torch.stack([big[m, :, y:y + 256, x:x + 256] for m in range(3) for x in range(0, 1024, 256)
                                   for y in range(0, 1024, 256)])
            feats = enc((patches - mean) / std)
        graph = {"x_imgN": feats[0:16], "x_imgA": feats[16:32], "x_imgL": feats[32:48], "x_cli": torch.randn(4, 1024),
                 "edge_index_imageN": FR.image_edge_index(), "edge_index_imageA": FR.image_edge_index(),
                 "edge_index_imageL": FR.image_edge_index(), "edge_index_cli": FR.cli_edge_index()}
        out = FR.fusion_forward(graph, state, mask=mask)
        loss = FR.fusion_loss([out], [mask], torch.tensor([1]))
        loss.backward()
        times.append(time.perf_counter() - t0)
    sec = sum(times[1:]) / len(times[1:])
    return {"value": 1.0 / sec, "unit": "patients/s", "cores": cores, "kind": "port", "cpu": cpu_model_name(),
            "sample": "%d timed single-patient steps (48 patches through torchvision resnet101 fp32 + oracle fusion head "
                      "fwd/objective/bwd), %.2f s each" % (reps, sec)}


# ----------------------------------------------------------------------------- our arm
def augment_leg(batch: int, size: int, peaks):
    """the training augmentation (dataloader.py:55-154) for one batch of VOC-sized decoded images.
    Host: drawing the decisions + building / packing the integer tables (one core; a DataLoader worker's share).
    Device: the four cvx_aug_* launches, with and without the host->device copy of the packed batch.  Beside it the
    unmodified reference's get_random_data + loader tail on one host core."""
    import numpy as np
    import torch
    from cervix_b200.utils import dataloader as D
    from oracle.ref_runner import augment_sources
    srcs = augment_sources(batch)
    np.random.seed(0)

    def pack():
        return D.pack_batch([(a, b, D.draw_params(b.shape[1], b.shape[0], (size, size))) for a, b in srcs], (size, size))

    pack()
    reps = 8
    plans, host_t = [], []
    for _ in range(reps):                     # fresh random sizes every time: the table cache mostly misses
        t0 = time.perf_counter()
        plans.append(pack().pin_memory())
        host_t.append((time.perf_counter() - t0) * 1e3)
    host_ms = sorted(host_t)[reps // 2]
    aug = D.DeviceAugmenter("cuda")
    for p in plans[:2]:
        aug.run(p)
    torch.cuda.synchronize()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # medians over the batches: the four launches of a batch are enqueued from Python, so a busy host core shows up as
    # idle gaps between them in a mean
    ms_e2e = sorted(timed(lambda: aug.run(p)) for p in plans)[reps // 2]
    blobs = [aug.upload(p) for p in plans]
    torch.cuda.synchronize()
    ms_dev = sorted(timed(lambda: aug.run(p, b)) for p, b in zip(plans, blobs))[reps // 2]
    src_bytes = sum(int(p.src.numel()) for p in plans) / reps
    out_bytes = batch * size * size * 4
    res = {"what": "get_random_data for a batch of %d decoded images (375x500 / 500x375) onto %dx%d canvases: bicubic + "
                   "nearest resize, flip, paste, blur (p=.25), rotation (p=.25), HSV jitter; uint8 in, uint8 out, bit-exact "
                   "with Pillow / OpenCV (tests/test_augment.py)" % (batch, size, size),
           "images_per_s_device": batch / (ms_dev * 1e-3), "ms_per_batch_device": ms_dev,
           "images_per_s_with_h2d": batch / (ms_e2e * 1e-3), "ms_per_batch_with_h2d": ms_e2e,
           "h2d_bytes_per_batch": int(src_bytes + sum(int(p.tables.numel()) * 4 + int(p.luts.numel()) + int(p.samples.numel()) for p in plans) / reps),
           "host_pack_ms_per_batch_one_core": host_ms, "launches_per_batch": 4,
           "roofline": {"bound": "hbm", "achieved": (src_bytes + out_bytes) / (ms_dev * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": (src_bytes + out_bytes) / (ms_dev * 1e-3) / 1e9 / peaks["hbm"],
                        "bytes_per_batch": int(src_bytes + out_bytes),
                        "note": "algorithmic bytes = decoded sources read once + augmented pixels and class maps written once; "
                                "the four launches also write and re-read two uint8 canvases (L2-resident at this size)"}}
    if reference_installed():
        r = ref_subprocess(["--augment", 64, "--size", size], timeout_s=300)
        if "error" not in r:
            res["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s", "cores": 1, "kind": "reference", "cpu": cpu_model_name(),
                                   "sample": "unmodified reference get_random_data + loader tail (dataloader.py:36-154) on %d of the same "
                                             "sources, one process (%.1f ms/image)" % (r["images"], r["ms_per_image"])}
        else:
            res["cpu_baseline"] = {"error": r["error"]}
    return res


def _time_launch(fn, flush, reps: int = 10):
    """Average duration of one launch: CUDA events on the launching (current) stream, L2 flushed between launches."""
    import torch
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def kernel_rooflines(batch: int, peaks):
    """Live per-launch rooflines of the kernels that dominate the step (ncu launch list: profiles/r02_ncu_launches_*):
      * tensor: the middle-flow pointwise GEMM 728 -> 728 @32x32 (150 launches per step = the time-dominant GEMM shape:
        forward with the BatchNorm-statistics epilogue, weight gradient), and the decoder 3x3 304 -> 256 @128x128 (the
        largest single launch) for contrast;
      * hbm: the fused depthwise backward of the same middle-flow tensor (the time-dominant bandwidth kernel).
    Algorithmic flops = 2*M*N*K; algorithmic bytes = each tensor the kernel must touch, once."""
    import torch
    from cervix_b200.backend import ConvGeom, get_backend
    B = get_backend()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    bf = lambda *shape: torch.randn(shape, device="cuda").bfloat16()   # noqa: E731
    out = {}
    # ---- middle-flow pointwise GEMM
    g = ConvGeom(batch, 32, 32, 728, 728, 1, 1, 1, 0, 1)
    x, dy, wp = bf(batch, 32, 32, 728), bf(batch, 32, 32, 728), bf(1, 728, 728)
    bias = torch.zeros(728, device="cuda")
    flops = 2.0 * batch * 32 * 32 * 728 * 728
    ms_f = _time_launch(lambda: B.conv_fwd_ex(x, wp, bias, g, want_stats=True), flush)
    ms_w = _time_launch(lambda: B.conv_wgrad(x, dy, g, True), flush)
    out["mid_pw_fwd"] = {"kernel": "conv_tc_fwd_2cta_kernel<1, 2> (middle-flow pointwise 728->728 @32x32, batch %d, BN statistics "
                                   "epilogue; 51 launches/step)" % batch, "ms_per_launch": ms_f, "flops_per_launch": flops,
                         "achieved": flops / (ms_f * 1e-3) / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s"}
    # DRAM traffic of one launch from the committed ncu --set full capture of this shape at batch 32
    # (profiles/r02_ncu_summary_v2.txt: dram__bytes_read.sum 48.8 MB + dram__bytes_write.sum 7.8 MB; the 47.7 MB output is
    # still in L2 when the kernel ends)
    out["mid_pw_fwd"]["traffic"] = 56.6e6 if batch == 32 else None
    out["mid_pw_fwd"]["traffic_source"] = "profiles/r02_ncu_summary_v2.txt (ncu --set full, batch 32): DRAM read 48.8 MB + write 7.8 MB per launch"
    out["mid_pw_wgrad"] = {"kernel": "conv_tc_wgrad_2cta_kernel (same shape; 51 launches/step)", "ms_per_launch": ms_w,
                           "flops_per_launch": flops, "achieved": flops / (ms_w * 1e-3) / 1e12, "peak": peaks["tf_burst"],
                           "unit": "TFLOP/s"}
    # ---- largest single launch
    g2 = ConvGeom(batch, 128, 128, 304, 256, 3, 3, 1, 1, 1)
    x2, wp2 = bf(batch, 128, 128, 304), bf(9, 256, 304)
    flops2 = 2.0 * batch * 128 * 128 * 256 * 304 * 9
    ms2 = _time_launch(lambda: B.conv_fwd(x2, wp2, None, g2, True), flush)
    out["cat_conv0_fwd"] = {"kernel": "conv_tc_fwd_2cta_kernel<0, 1> (decoder 3x3 304->256 @128x128, batch %d)" % batch,
                            "ms_per_launch": ms2, "flops_per_launch": flops2, "achieved": flops2 / (ms2 * 1e-3) / 1e12,
                            "peak": peaks["tf_burst"], "unit": "TFLOP/s"}
    for v in out.values():
        v["frac"] = v["achieved"] / v["peak"]
        v["peak_source"] = peaks["source"] + " burst bf16 (kernel timed alone)"
    del x2, wp2
    # ---- dominant bandwidth kernel: fused depthwise backward (data + weight gradient + BatchNorm backward sums)
    gd = ConvGeom(batch, 32, 32, 728, 728, 3, 3, 1, 1, 1)
    w9c = torch.randn(9, 728, device="cuda")
    sc, sh = torch.rand(728, device="cuda") + 0.5, torch.randn(728, device="cuda") * 0.1
    ms_d = _time_launch(lambda: B.dwf_bwd(dy, None, None, None, x, w9c, sc, sh, True, None, gd, True), flush)
    nbytes = 3.0 * batch * 32 * 32 * 728 * 2          # read dd, read x, write g (bf16)
    hbm = {"bound": "hbm", "kernel": "dwf_bwd_kernel<true, false> (depthwise 3x3 backward of the middle-flow tensor [%d,32,32,728]: "
                                      "data gradient + weight gradient + previous BatchNorm's backward sums; 37 launches/step)" % batch,
           "achieved": nbytes / (ms_d * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s", "bytes_per_launch": nbytes,
           "ms_per_launch": ms_d, "peak_source": peaks["source"] + " HBM copy bandwidth",
           "traffic": 113.9e6 if batch == 32 else None,
           "traffic_source": "profiles/r02_ncu_summary_v2.txt (ncu --set full of this launch at batch 32): DRAM read 95.5 MB + write "
                             "18.4 MB - below the algorithmic 143 MB because the gradient it consumes and most of what it "
                             "writes stay in L2"}
    hbm["frac"] = hbm["achieved"] / hbm["peak"]
    return out, hbm


def fit_one_epoch_throughput(model, imgs_h, pngs_h, labels_h, steps: int):
    """Images/s of the drop-in ``fit_one_epoch`` on the same model: torch.optim.Adam handed in as the reference's train.py
    does, pinned host batches, fp16=True (-> bf16 engine).  Two epochs: the first captures the graph, the second is timed."""
    import contextlib
    import io
    import tempfile
    import numpy as np
    import torch
    from cervix_b200.utils.utils_fit import fit_one_epoch

    class _Hist:
        val_loss: list = []

        def append_loss(self, *a):
            pass

    class _Ev:
        def on_epoch_end(self, *a):
            pass

    opt = torch.optim.Adam(model.parameters(), 1e-4, betas=(0.9, 0.999), weight_decay=0)
    cls_w = np.array(CLS_WEIGHTS, np.float32)
    batch = (imgs_h, pngs_h, labels_h)
    val = [(imgs_h[:2], pngs_h[:2], labels_h[:2])]
    stamps = {}
    first = 6       # the graph is captured while the third batch is being stepped (~1 s of host time): start the clock at
                    # the seventh pull, when that is safely over

    def gen(n):
        # fit_one_epoch pulls batch i one step ahead of issuing step i-1 (BatchPrefetcher) and pulls one element past
        # epoch_step to see the end; a device sync at pull a and at pull b therefore brackets exactly b - a steps
        for i in range(n + 1):
            if i == first or i == n:
                torch.cuda.synchronize()
                stamps[i] = time.perf_counter()
            yield batch

    with tempfile.TemporaryDirectory() as save_dir, contextlib.redirect_stdout(io.StringIO()), \
            contextlib.redirect_stderr(io.StringIO()):
        n = steps + first
        fit_one_epoch(model, model, _Hist(), _Ev(), opt, 0, n, 1, gen(n), val, 3, True, True, True, cls_w, 5, True, None, 1000,
                      save_dir, 0)
    sec = (stamps[n] - stamps[first]) / (n - first)
    tr = getattr(model, "_cvx_trainer", None)
    engine = "SegTrainer (CUDA graph)" if tr is not None and tr.graph is not None else "eager autograd + optimizer.step()"
    if tr is not None:
        tr.close()
        model._cvx_trainer = None
    return {"value": imgs_h.shape[0] / sec, "unit": UNIT, "steps": n - first, "ms_per_step": sec * 1e3,
            "engine": engine,
            "what": "utils.utils_fit.fit_one_epoch(model, torch.optim.Adam, ...) on pinned host batches (fp32 images, int64 "
                    "class maps, fp32 one-hot labels as the reference's loader yields them), device-synchronised wall clock "
                    "over the steady-state steps of the training phase"}


def strong_scaling_leg(model, bsz: int, size: int, world: int, rank: int, steps: int, comm_in_graph: bool):
    import torch
    import torch.distributed as dist
    from cervix_b200.engine import SegTrainer
    trainer = SegTrainer(model, lr=1e-4, betas=(0.9, 0.999), cls_weights=CLS_WEIGHTS, num_classes=5, world_size=world)
    imgs, pngs, _ = synthetic_batch(bsz, size, seed=100 + rank)
    imgs, pngs = imgs.cuda(), pngs.cuda()
    split = os.environ.get("CERVIX_SPLIT_BACKWARD", "1") != "0" and not comm_in_graph
    if not (split and trainer.capture_split(imgs, pngs, None) is not None):
        trainer.capture(imgs, pngs, None, comm_in_graph=comm_in_graph)
    for _ in range(3):
        trainer.step_graphed(imgs, pngs)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        trainer.step_graphed(imgs, pngs)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    trainer.close()
    del trainer
    return {"value": world * bsz / (ms * 1e-3), "unit": UNIT, "global_batch": world * bsz, "per_gpu_batch": bsz, "ms_per_step": ms,
            "scaling": "strong (BASELINE configs[3]: global batch 256)"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as entry
    if not os.path.exists(entry.LIB):
        if rank == 0:
            entry.build()
        if world > 1:
            dist.barrier()
    from cervix_b200.backend import get_backend
    from cervix_b200.engine import SegTrainer
    from cervix_b200.nets.deeplabv3_plus import DeepLab

    B = get_backend()
    peaks = load_peaks()
    bsz, size = args.batch, args.size
    torch.manual_seed(0 + rank)
    model = DeepLab(5, "xception", False, 16).set_compute_dtype(torch.bfloat16)
    from cervix_b200.nets.deeplabv3_training import weights_init
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        weights_init(model)   # the reference's init for a non-pretrained run (train.py:314-315)
    model.cuda().train()
    if world > 1:   # identical initial weights on every rank, as DDP's broadcast would give
        for p in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(p.data, 0)
    trainer = SegTrainer(model, lr=1e-4, betas=(0.9, 0.999), cls_weights=CLS_WEIGHTS, num_classes=5, world_size=world)

    imgs_h, pngs_h, labels_h = synthetic_batch(bsz, size, seed=rank)   # the product arm never touches oracle/
    imgs_h, pngs_h, labels_h = imgs_h.pin_memory(), pngs_h.pin_memory(), labels_h.pin_memory()
    imgs, pngs, labels = imgs_h.cuda(), pngs_h.cuda(), labels_h.cuda()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    use_graph = not args.no_graph
    comm_in_graph = os.environ.get("CERVIX_COMM_IN_GRAPH", "0") == "1"
    split = world > 1 and os.environ.get("CERVIX_SPLIT_BACKWARD", "1") != "0" and not comm_in_graph

    def capture_step(a, b):
        # data parallel: two graphs (backward split at the end of the entry flow) with the big all-reduce between them
        if not (split and trainer.capture_split(a, b, None) is not None):
            trainer.capture(a, b, None, comm_in_graph=comm_in_graph)

    used_split = False
    if use_graph:
        capture_step(imgs, pngs)
        used_split = bool(getattr(trainer, "_split", False))
        step_fn = lambda a, b, c=None: trainer.step_graphed(a, b)   # noqa: E731
    else:
        step_fn = lambda a, b, c=None: trainer.step(a, b, c)        # noqa: E731

    # ---- device-resident throughput ------------------------------------------------------
    for _ in range(args.warmup):
        step_fn(imgs, pngs, labels)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = B.lib.cvx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_fn(imgs, pngs, labels)
    e1.record()
    sync_all()
    launches = B.lib.cvx_launch_count() - launches0
    if use_graph:   # replays do not pass through the library's launch counter: count one eager step instead
        c0 = B.lib.cvx_launch_count()
        trainer.step(imgs, pngs, None)
        launches = (B.lib.cvx_launch_count() - c0) * args.steps
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    losses = [float(v) for v in res.cpu()]

    # ---- end to end: every step's batch comes from pinned HOST memory, prefetched one step ahead on a copy stream
    # (two device buffer sets, so two copies can be in flight); the 4 loss/metric scalars are read back to the host every
    # step.  Two input contracts are timed:
    #   "u8"   what a loader decodes: uint8 [B,H,W,3] pixels + uint8 class maps (dataloader.py:40-42 run on the device by
    #          cvx_finish_batch_u8) - 1 MB per image; this is the headline e2e
    #   "f32"  the reference DataLoader's tensors as they arrive in fit_one_epoch (utils_fit.py:52-58): fp32 NCHW images +
    #          int64 class maps - 5.2 MB per image
    from cervix_b200.engine import BatchPrefetcher

    def e2e_run(host_batch, k):
        def host_batches(n):
            for _ in range(n):
                yield host_batch
        # ONE prefetcher for warm-up and timed steps: building it allocates its two device buffer sets (cudaMalloc behind
        # two resident graph pools cost 60-150 ms, once per epoch in a real run), which is not a per-step cost.  Every
        # timed step still has exactly one host->device copy enqueued inside the timed region (that of the NEXT batch).
        warm = 6
        pf = iter(BatchPrefetcher(host_batches(warm + k + 1)))
        for _ in range(warm):
            bi, bp = next(pf)
            step_fn(bi, bp, None).cpu()
        sync_all()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e2.record()
        prev = None
        marks = []
        for _ in range(k):
            bi, bp = next(pf)
            r = step_fn(bi, bp, None)
            if use_graph:
                r = r.clone()                             # the graph's output buffer is overwritten by the next replay
            if prev is not None:
                prev.cpu()                                # result of the previous step (keeps 1 step in flight)
                marks.append(time.perf_counter())
            prev = r
        prev.cpu()
        marks.append(time.perf_counter())
        e3.record()
        sync_all()
        gaps = sorted((b - a) * 1e3 for a, b in zip(marks, marks[1:]))
        e2e_step_stats.append({"median_ms": gaps[len(gaps) // 2], "max_ms": gaps[-1], "first_result_ms": (marks[0] - t0) * 1e3})
        return max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)

    e2e_step_stats = []

    # the e2e window is at least 24 steps (per-step gaps are reported beside the mean: `step_gaps`)
    e2e_steps = max(args.steps, 24)
    ms_e2e_f32 = e2e_run((imgs_h, pngs_h), e2e_steps)
    imgs_u8_h = (imgs_h.permute(0, 2, 3, 1) * 255.0).round().to(torch.uint8).contiguous().pin_memory()
    pngs_u8_h = pngs_h.to(torch.uint8).pin_memory()
    if use_graph:
        capture_step(imgs_u8_h.cuda(), pngs_u8_h.cuda())      # same step over uint8 static input buffers
    ms_e2e = e2e_run((imgs_u8_h, pngs_u8_h), e2e_steps)

    t = torch.tensor([ms_total, ms_e2e, ms_e2e_f32], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_e2e_f32 = float(t[0]), float(t[1]), float(t[2])
    # ---- the reference's own entry point: fit_one_epoch(model, torch.optim.Adam, pinned host batches in the loader's
    # contract: fp32 images, int64 class maps, fp32 one-hot labels) - the drop-in routes it onto the same engine
    # (utils_fit.py:31-198; single GPU only: under torchrun the scripts wrap the model in DDP first)
    fit_ips = None
    trainer.close()
    del step_fn, trainer
    torch.cuda.empty_cache()
    if world == 1 and not args.no_fit:
        fit_ips = fit_one_epoch_throughput(model, imgs_h, pngs_h, labels_h, max(args.steps, 12))
    # ---- BASELINE configs[3] as written: GLOBAL batch 256 split over the ranks (train.py:499 `batch_size // ngpus`);
    # the main line is weak scaling at 32 images per GPU (identical at N = 8).  N = 4 -> 64 images per GPU is timed here;
    # N = 2 -> 128 per GPU needs ~140 GB of saved activations and is not attempted.
    strong = None
    strong_bsz = int(os.environ.get("CERVIX_BENCH_STRONG_BSZ", "0")) or (256 // world if 256 % world == 0 else 0)
    if world > 1 and 32 < strong_bsz <= 64 and not args.no_strong:
        free_gb = torch.tensor([torch.cuda.mem_get_info()[0] / 2 ** 30], device="cuda")
        dist.all_reduce(free_gb, op=dist.ReduceOp.MIN)          # every rank takes the same decision
        if float(free_gb) < 110.0:
            strong = {"skipped": "%.0f GB free on the fullest GPU; a %d-image step keeps ~95 GB of activations" % (float(free_gb), strong_bsz)}
        else:
            strong = strong_scaling_leg(model, strong_bsz, size, world, rank, args.steps, comm_in_graph)
    classifier = None
    if not args.no_classifier:      # every rank takes part (patients are sharded, head gradients all-reduced)
        del model
        torch.cuda.empty_cache()
        classifier = classifier_throughput(max(args.steps, 20), 3, patients=16 if world == 1 else 64,
                                           world=world, rank=rank,
                                           variants=(("imgN", "imgA", "imgL", "cli"), ("imgN", "imgA", "imgL"), ("imgN", "imgL")),
                                           use_graph=use_graph)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_total / args.steps
    value = world * bsz * args.steps / (ms_total * 1e-3)
    e2e_value = world * bsz * e2e_steps / (ms_e2e * 1e-3)
    e2e_f32_value = world * bsz * e2e_steps / (ms_e2e_f32 * 1e-3)
    kernels, roof_hbm = kernel_rooflines(bsz, peaks)
    step_tf = value / world * TRAIN_GFLOP_PER_IMAGE / 1e3
    # `roofline` describes the STEP (what the metric measures): convolution flops of one training step / step time against
    # the sustained bf16 peak.  `dominant_kernel` is the time-dominant GEMM shape of the committed launch list, timed live;
    # `roofline_hbm` is the time-dominant bandwidth kernel.
    roof = {"bound": "tensor", "scope": "whole training step (all kernels, GEMM and bandwidth-bound alike)",
            "achieved": step_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": step_tf / peaks["tf_sustained"],
            "flops_per_step": TRAIN_GFLOP_PER_IMAGE * 1e9 * bsz, "ms_per_step": ms_step,
            "peak_source": peaks["source"] + " sustained bf16 (kernel timed inside a long step)",
            "traffic": None, "kernel": kernels["mid_pw_fwd"]["kernel"], "dominant_kernel": kernels["mid_pw_fwd"],
            "other_kernels": {k: v for k, v in kernels.items() if k != "mid_pw_fwd"}}
    cpu_baseline = torch_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        cb = cpu_reference_baseline(4, 1, budget_s=25.0)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "cpu", "sample")}
    if world == 1 and not args.no_gpu_baseline:
        torch.cuda.empty_cache()
        torch_gpu = torch_gpu_baseline(bsz, size)
        for k in ("fp16_autocast", "fp16_autocast_channels_last"):
            if isinstance(torch_gpu.get(k), dict) and "images_per_s" in torch_gpu[k]:
                torch_gpu[k]["ours_over_this"] = value / torch_gpu[k]["images_per_s"]
                torch_gpu[k]["ours_e2e_over_this"] = e2e_value / torch_gpu[k]["images_per_s"]
    augment = None
    if world == 1 and not args.no_augment:
        try:
            augment = augment_leg(bsz, size, peaks)
        except Exception as e_aug:      # report, never fail the bench line
            augment = {"error": repr(e_aug)}
    if classifier is not None and world == 1 and not args.no_cpu_baseline:
        try:
            classifier["cpu_baseline"] = classifier_cpu_baseline()
        except Exception as e:   # torchvision missing etc.: report, never fail the bench line
            classifier["cpu_baseline"] = {"error": repr(e)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "DeepLabv3+ Xception ds=16 bf16 training (fwd + focal+dice + bwd + Adam), "
                               "batch %d per GPU at %dx%d, 5 classes (BASELINE configs[2]/[3])" % (bsz, size, size),
                   "global_batch": world * bsz, "per_gpu_batch": bsz, "parallelism": "dp%d" % world,
                   "cuda_graph": bool(use_graph), "allreduce": (None if world == 1 else (
                       "bucketed NCCL all-reduce forked from the gradient hooks, captured inside the step's CUDA graph"
                       if (use_graph and comm_in_graph) else "bucketed NCCL all-reduce overlapped with backward" if not use_graph
                       else "two CUDA graphs per step: the all-reduce of 97 % of the gradient (everything behind the entry flow) "
                            "runs on NCCL's stream under the second graph (the entry flow's backward)" if used_split
                       else "bucketed NCCL all-reduce after the graph replay, optimizer of bucket k under the all-reduce of bucket k+1")),
                   "l2": "per-step working set (tens of GB of activations) far exceeds the 126 MB L2"},
        "roofline": roof,
        "roofline_hbm": roof_hbm,
        "cpu_baseline": cpu_baseline,
        "torch_gpu_baseline": torch_gpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": imgs_u8_h.numel() + pngs_u8_h.numel(),
                "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                "inputs": "pinned host uint8 [B,H,W,3] pixels + uint8 class maps (what a loader decodes)",
                "step_gaps": e2e_step_stats[1] if len(e2e_step_stats) > 1 else None,
                "fp32_contract": {"value": e2e_f32_value, "h2d_bytes_per_step": imgs_h.numel() * 4 + pngs_h.numel() * 8,
                                  "ms_per_step": ms_e2e_f32 / e2e_steps, "step_gaps": e2e_step_stats[0] if e2e_step_stats else None,
                                  "inputs": "pinned host fp32 NCHW images + int64 class maps (utils_fit.py:52-58)"}},
        "fit_one_epoch": fit_ips,
        "augment": augment,
        "configs3_global_batch_256": strong,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "classifier": classifier,
        "losses_last_step": {"ce": losses[0], "focal": losses[1], "dice": losses[2], "f_score": losses[3]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch run of the reference on this GPU")
    ap.add_argument("--no-classifier", action="store_true", help="skip the severity-classifier leg")
    ap.add_argument("--no-graph", action="store_true", help="run the single-GPU step eagerly instead of as a CUDA graph")
    ap.add_argument("--no-fit", action="store_true", help="skip the fit_one_epoch (reference entry point) leg")
    ap.add_argument("--no-augment", action="store_true", help="skip the device-augmentation leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the configs[3] global-batch-256 leg at N = 4")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
