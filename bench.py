#!/usr/bin/env python
"""Headline benchmark: U-Net images/sec (BASELINE.json `metric`), configs[1] per GPU:
fixture best_unet_model.pth, batch 64 x 3 x 512 x 512, bf16 tensor-core forward on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU forward (oracle port)

One process per GPU (torchrun for N > 1); images are independent, so ranks shard the batch
with no data-path collective ("scaling": "weak": 64 images per GPU per step).  Rank 0 prints
ONE JSON line.  See DESIGN.md "Measurement" for what each field means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "unet_forward_images_per_sec"
UNIT = "images/s"
GFLOP_PER_IMAGE_512 = 385.406          # algorithmic FLOPs of the 23 convs at 512x512


def layer_flops(layer, h, w) -> float:
    """Algorithmic FLOPs per image of one row of the layer table (padding taps counted)."""
    hh, ww = h >> layer.level, w >> layer.level
    if layer.kind in (0, 1):       # 3x3 convs (stem, conv3x3)
        return 2.0 * hh * ww * layer.cout * layer.cin * 9
    if layer.kind == 2:            # convT 2x2 s2: 4 output pixels per input pixel
        return 2.0 * hh * ww * layer.cin * layer.cout * 4
    return 2.0 * hh * ww * layer.cin * layer.cout      # 1x1 head


def layer_bytes(layer, h, w, n_classes=3) -> float:
    """Algorithmic HBM bytes per image of one layer in the fused plan: bf16 NHWC
    activations read once per consumer and written once, pool as a second output, concat virtual,
    out_conv + threshold fused into conv1.net.3 (fp32 logits + u8 mask out); weights excluded."""
    px = (h >> layer.level) * (w >> layer.level)
    name = layer.name.decode()
    if layer.kind == 0:                                   # stem: fp32 NCHW in, bf16 out
        return px * (layer.cin * 4 + layer.cout * 2)
    if layer.kind == 2:                                   # convT: 4 output pixels per input pixel
        return px * layer.cin * 2 + 4 * px * layer.cout * 2
    if layer.kind == 3:
        return 0.0
    b = px * layer.cin * 2
    if name == "conv1.net.3":
        return b + px * n_classes * (4 + 1)
    b += px * layer.cout * 2
    if name.startswith("down") and name.endswith(".3"):
        b += px // 4 * layer.cout * 2                     # fused 2x2 max-pool output
    return b


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"tflops": float(p["bf16_tflops_sustained"]), "tflops_burst": float(p["bf16_tflops"]),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "MEASURED_PEAKS.json (sustained bf16 figure: kernels timed inside a long step)"}
    except Exception:
        return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
                "source": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


def ncu_traffic_bytes():
    """DRAM bytes (read + write) of the 21 tensor-core launches of one step, from the newest committed
    `ncu --set full` capture (profiles/*_ncu_raw.csv, batch 64 @ 512x512); None if absent."""
    import csv
    import glob
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", "*_ncu_raw.csv")) if "enhance" not in f)
    if not files:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        h, unit, data = rows[0], rows[1], rows[2:]
        ir, iw, ik = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("Kernel Name")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit[ir]]
        tc = [r for r in data if "conv_tc_kernel" in r[ik] and "conv_tc_kernel<64, 1, 3," not in r[ik]
              and "(int)64, (int)1, (int)3" not in r[ik]]            # all launches but the stem (A_STEM)
        if len(tc) != 21:
            return None, None
        return sum(float(r[ir]) + float(r[iw]) for r in tc) * scale, os.path.basename(files[-1])
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_forward_rate(state, n_images: int, warmup: int, batch: int = 1):
    """images/s of the oracle's fp32 CPU forward (the reference's algorithm on the host cores)."""
    import torch
    from oracle.unet_oracle import oracle_forward
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x = synthetic_invoices(batch, 512, 512, seed=42)
    for _ in range(warmup):
        oracle_forward(state, x)
    times = []
    for _ in range(max(1, n_images // batch)):
        t0 = time.perf_counter()
        oracle_forward(state, x)
        times.append(time.perf_counter() - t0)
    return batch / statistics.median(times), cores, times


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path.  The reference is
    pure Python/PyTorch and cannot travel to the GPU box, so this times the oracle port (the same
    torch.nn.functional CPU kernels, called functionally) on all host cores; rank 0 only."""
    if rank != 0:
        return
    from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state
    state = make_fixture_state()
    per_step = 2                                     # bounded sample: 2 of the 64 images per step
    rate, cores, times = cpu_forward_rate(state, per_step * args.steps, max(1, args.warmup), batch=per_step)
    ms = 1e3 * statistics.median(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] shape on the CPU path: fixture best_unet_model.pth, 3x512x512 "
                               f"invoices, fp32 torch CPU forward, {per_step} images per step (bounded sample "
                               "of the batch-64 step)", "image": "3x512x512", "batch_per_step": per_step},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {per_step} images, median step"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_JSON_FD = None


def _reserve_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries print there too (NCCL's version banner
    with NCCL_DEBUG set, cuDNN / driver notices), so file descriptor 1 is pointed at stderr for the whole
    run and the JSON line is written to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the informational sections (batch-1 latency, run_unet, crop enhancement): profiling runs")
    ap.add_argument("--layers-out", default=None, help="write the per-layer table (JSON) here")
    ap.add_argument("--graph", type=int, default=0, help="1: replay the step as a CUDA graph")
    ap.add_argument("--e2e-chunk", type=int, default=64,
                    help="images per forward inside the end-to-end call (copies of chunk i+1 overlap chunk i)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from tw_invoice_unet_ocr_llm_b200 import _native as nat
    from tw_invoice_unet_ocr_llm_b200.launcher import GpuWorker
    from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices_u8

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA (B200) device: there is no CPU path for --impl b200")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_cores = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # one process per GPU: keep this rank and its pinned staging on the cores / memory local to its GPU
        from tw_invoice_unet_ocr_llm_b200.launcher import bind_to_gpu_numa
        numa_cores = bind_to_gpu_numa(local_rank)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version banner must not land on stdout
        dist.init_process_group("nccl", device_id=dev)

    B, S = args.batch, args.size
    state = make_fixture_state()
    worker = GpuWorker(state, dev, chunk=min(B, args.e2e_chunk))   # public multi-GPU launcher building block
    eng = worker.engine
    thr = [0.25, 0.40, 0.30]

    # synthetic invoice-shaped batch, distinct per rank; 8 distinct frames tiled to B
    base = synthetic_invoices_u8(min(8, B), S, S, seed=7 + rank)
    frames_u8 = np.concatenate([base] * ((B + len(base) - 1) // len(base)))[:B]
    x_dev = torch.from_numpy(frames_u8.astype(np.float32) / 255.0).permute(0, 3, 1, 2).contiguous().to(dev)
    logits = torch.empty((B, 3, S, S), dtype=torch.float32, device=dev)
    mask = torch.empty((B, 3, S, S), dtype=torch.uint8, device=dev)

    def step():
        eng.run(x_dev, want_logits=True, thresholds=thr, logits_out=logits, mask_out=mask)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, k):
        """K calls of fn bracketed by barrier + synchronize; device time via CUDA events on the
        launching (current) stream; max over ranks."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- device-resident throughput (`value`)
    for _ in range(args.warmup):
        step()
    if args.graph:
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(dev)
        with torch.cuda.stream(cap):
            step()
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(g, stream=cap):
                step()
        eager_step = step
        step = g.replay
        for _ in range(args.warmup):
            step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches_per_step = eng.last_launch_count()
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---------------- per-layer pass for the roofline (same K steps, events between launches)
    if args.graph:
        step = eager_step
    eng.set_option("profile", 1)
    step()
    acc = None
    sync_all()
    for _ in range(args.steps):
        step()
        t = eng.layer_times_ms()
        acc = t if acc is None else [a + b for a, b in zip(acc, t)]
    eng.set_option("profile", 0)
    layer_ms = [a / args.steps for a in acc]
    layers = eng.layers
    peaks = read_peaks()
    table, tc_flops, tc_ms, tc_bytes = [], 0.0, 0.0, 0.0
    for i, l in enumerate(layers):
        if l.kind == nat.HEAD:
            continue                                  # fused into conv1.net.3
        fl = layer_flops(l, S, S) * B
        if l.name.decode() == "conv1.net.3":
            fl += layer_flops(layers[-1], S, S) * B
        ms = layer_ms[i]
        tf = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        tensor = l.kind in (nat.CONV3X3, nat.CONVT2X2)
        by = layer_bytes(l, S, S) * B + l.w_bytes
        if tensor:
            tc_flops += fl
            tc_ms += ms
            tc_bytes += by
        table.append({"layer": l.name.decode(), "kernel": "conv_tc_kernel (tcgen05)" if tensor else
                      ("conv_tc_kernel<A_STEM> (tcgen05, in-kernel im2col; HBM / issue bound)"
                       if eng.get_option("stem_tc") else "stem_conv_kernel (CUDA cores)"),
                      "ms": round(ms, 4), "gflop": round(fl / 1e9, 2), "tflops": round(tf, 1),
                      "algorithmic_gb": round(by / 1e9, 3), "gb_per_s": round(by / 1e9 / (ms * 1e-3), 1) if ms > 0 else 0.0,
                      "frac_of_peak": round(tf / peaks["tflops"], 3)})
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic_bytes() if (B == 64 and S == 512) else (None, None)
    roofline = {
        "bound": "tensor", "kernel": "conv_tc_kernel<BN,TAPS,AMODE,EPI> (21 launches per step: 17 conv3x3 + 4 convT)",
        "achieved": round(achieved, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
        "frac": round(achieved / peaks["tflops"], 4), "traffic": traffic,
        "traffic_note": (f"DRAM read+write bytes of the same 21 launches of one step, ncu --set full ({traffic_src}); "
                         f"algorithmic bytes of those launches: {tc_bytes / 1e9:.2f} GB per step") if traffic else None,
        "peak_source": peaks["source"],
        "how": "sum of algorithmic conv FLOPs of the 21 tcgen05 launches / sum of their CUDA-event durations "
               "(events recorded between launches on the launching stream, averaged over the K steps)",
        "share_of_step": round(tc_ms / sum(layer_ms), 4),
    }
    if args.layers_out and rank == 0:
        with open(args.layers_out, "w") as f:
            json.dump({"batch": B, "size": S, "layers": table, "ms_per_step_profiled": sum(layer_ms)}, f, indent=1)

    # ---------------- end to end through the public launcher API: pinned host frames in, host masks out
    frames_host = torch.from_numpy(frames_u8).pin_memory()
    masks_host = torch.empty((B, 3, S, S), dtype=torch.uint8).pin_memory()

    # Steady-state serving loop: every step uploads its own 64 frames and downloads its own masks
    # (all inside the timed region); the launcher double-buffers, so step k+1's upload overlaps
    # step k's forward.  Alternating host mask buffers stand in for the consumer of step k.
    masks_alt = torch.empty((B, 3, S, S), dtype=torch.uint8).pin_memory()

    def e2e_run(k):
        for i in range(k):
            worker.segment_async(frames_host, masks_host if i % 2 == 0 else masks_alt)
        worker.join_current_stream()

    e2e_run(args.warmup)
    worker.synchronize()
    e2e_ms = timed(lambda: e2e_run(args.steps), 1) / args.steps
    e2e = {"value": world * B / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(frames_host.numel()), "d2h_bytes_per_step": int(masks_host.numel()),
           "chunk": worker.chunk,
           "api": "tw_invoice_unet_ocr_llm_b200.launcher.GpuWorker.segment_async + synchronize (pinned uint8 frames -> "
                  "pinned uint8 masks, double-buffered across steps)"}

    # ---------------- batch-1 latency (BASELINE.json configs[4]): one resident 3x512x512 frame ->
    # logits + masks, synchronised per call; p50/p95 over 200 calls (rank 0, informational)
    lat = None
    if rank == 0 and not args.no_extras:
        x1 = x_dev[:1].contiguous()
        l1 = torch.empty((1, 3, S, S), dtype=torch.float32, device=dev)
        m1 = torch.empty((1, 3, S, S), dtype=torch.uint8, device=dev)
        ts = []
        for i in range(220):
            t0 = time.perf_counter()
            eng.run(x1, want_logits=True, thresholds=thr, logits_out=l1, mask_out=m1)
            torch.cuda.synchronize(dev)
            if i >= 20:
                ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        lat = {"p50_ms": ts[len(ts) // 2], "p95_ms": ts[int(len(ts) * 0.95)], "calls": len(ts),
               "what": "engine.run on one resident 3x512x512 frame, host-synchronised per call"}
        # the same forward replayed as a CUDA graph
        try:
            g1 = torch.cuda.CUDAGraph()
            cap = torch.cuda.Stream(dev)
            with torch.cuda.stream(cap):
                eng.run(x1, want_logits=True, thresholds=thr, logits_out=l1, mask_out=m1)
                torch.cuda.synchronize(dev)
                with torch.cuda.graph(g1, stream=cap):
                    eng.run(x1, want_logits=True, thresholds=thr, logits_out=l1, mask_out=m1)
            ts = []
            for i in range(220):
                t0 = time.perf_counter()
                g1.replay()
                torch.cuda.synchronize(dev)
                if i >= 20:
                    ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            lat["graph_replay"] = {"p50_ms": ts[len(ts) // 2], "p95_ms": ts[int(len(ts) * 0.95)]}
            del g1
        except Exception as e:          # informational only
            lat["graph_replay"] = {"error": repr(e)}

    # ---------------- run_unet end to end (BASELINE.json configs[4]b): 1920x1080 RGB frame -> masks + crops
    # through the reference-facing entry point with the model cached; GPU resize vs host PIL resize
    if rank == 0 and not args.no_extras:
        try:
            import tempfile
            from PIL import Image
            from tw_invoice_unet_ocr_llm_b200 import inference as inf
            with tempfile.TemporaryDirectory() as d:
                ckpt = os.path.join(d, "best_unet_model.pth")
                torch.save(state, ckpt)
                rgb = Image.fromarray(synthetic_invoices_u8(1, 1080, 1920, seed=11)[0])
                rgba = rgb.convert("RGBA")
                res = {}
                for tag, im in (("gpu_resize", rgb), ("host_pil_resize", rgba)):
                    ts = []
                    for i in range(60):
                        t0 = time.perf_counter()
                        inf.run_unet(im, ckpt)
                        if i >= 10:
                            ts.append((time.perf_counter() - t0) * 1e3)
                    ts.sort()
                    res[tag] = {"p50_ms": ts[len(ts) // 2], "p95_ms": ts[int(len(ts) * 0.95)]}
                # + the crop enhancement the app applies next (app_camera.py:800-811), same device frame
                ts = []
                for i in range(60):
                    t0 = time.perf_counter()
                    inf.run_unet_enhanced(rgb, ckpt)
                    if i >= 10:
                        ts.append((time.perf_counter() - t0) * 1e3)
                ts.sort()
                res["with_crop_enhancement"] = {"p50_ms": ts[len(ts) // 2], "p95_ms": ts[int(len(ts) * 0.95)]}
                # the same sequence with the reference's cv2 calls on the host for the enhancement
                # (this fixture's crops cover most of the frame: ~90 Mpixel of enhanced output per call)
                try:
                    sys.path.insert(0, os.path.join(ROOT, "tools"))
                    import enhance_bench
                    ts = []
                    for i in range(6):
                        t0 = time.perf_counter()
                        _, crops = inf.run_unet(rgb, ckpt)
                        out_px = 0
                        for f, c in crops.items():
                            if c is not None:
                                e = enhance_bench._cv2_chain(np.array(c.convert("RGB")), inf.ENHANCE_KINDS[f])
                                out_px += e.size
                        if i >= 1:
                            ts.append((time.perf_counter() - t0) * 1e3)
                    ts.sort()
                    res["with_crop_enhancement_cv2_on_host"] = {"p50_ms": ts[len(ts) // 2], "enhanced_megapixels": out_px / 1e6}
                except Exception as e:
                    res["with_crop_enhancement_cv2_on_host"] = {"error": repr(e)}
            if lat is not None:
                lat["run_unet_1080p"] = res
        except Exception as e:          # informational only
            if lat is not None:
                lat["run_unet_1080p"] = {"error": repr(e)}

    # ---------------- OCR crop enhancement (rank 0, informational): 192 ragged crops
    # through enhance_batch vs the reference's cv2 calls on the host cores
    enh = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import enhance_bench
            enh = enhance_bench.measure(quick=True)
        except Exception as e:          # informational only
            enh = {"error": repr(e)}

    # ---------------- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, times = cpu_forward_rate(state, 8, 2, batch=1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{len(times)} single-image 3x512x512 fp32 forwards of the oracle after 2 warm-ups, median "
                         f"{statistics.median(times):.3f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"configs[1]: fixture best_unet_model.pth (seeded, same 136-key fp32 format), "
                                   f"batch {B} x 3x{S}x{S} synthetic invoices per GPU, bf16 NHWC tensor-core forward, "
                                   "fp32 NCHW input resident in HBM, outputs fp32 logits + uint8 masks",
                       "numa_bound_cores_rank0": len(numa_cores) if numa_cores else None,
                       "batch_per_gpu": B, "image": f"3x{S}x{S}", "parallelism": f"dp{world} (batch sharding, no collective)",
                       "l2": "no flush needed: per-step working set (~9 GB of activations, 201 MB input) exceeds the 126 MB L2",
                       "gflop_per_image": GFLOP_PER_IMAGE_512 * (S * S) / (512 * 512),
                       "achieved_tflops_whole_step": value / world * GFLOP_PER_IMAGE_512 * (S * S) / (512 * 512) / 1e3},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "latency_b1": lat, "enhance": enh, "clocks": clocks,
            "gpu_launches": launches_per_step * args.steps,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
