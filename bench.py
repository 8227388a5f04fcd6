#!/usr/bin/env python3
"""bench.py -- the BC-linear hot path on B200, measured the way BASELINE.json quotes it.

Workload (config.workload = "llama2-7b-decode-b128", BASELINE.json configs[1]): one decode step of batch 1 through
ALL 7 block-circulant projections (q,k,v,o 4096->4096; gate,up 4096->11008; down 11008->4096; B=128) of ALL 32
layers of a Llama-2-7B-shaped model: 224 BC linears, 224 distinct random-init weight spectra (207 MB fp32 -- larger
than the 126 MB L2, and two such sets alternate between steps), synthetic fp16 activations.
A step is ONE launch of the grouped stream kernel (bcl_forward_grouped -> bcl::ws::stream_ws_kernel) over those 224
independent problems.  `detail` re-times the same 224 projections at the launch granularities a real decoder has
(one grouped launch per layer, one launch per module; CUDA-graph replay) so nothing hides behind the grouping.

    value            tokens/s, whole job (N GPUs decode N independent streams: data parallel, "weak")
    roofline         algorithmic bytes of a step / CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
    e2e              same step through the host-buffer entry point: pinned x -> H2D -> launch -> D2H -> sync
    cpu_baseline     the reference's CPU torch.fft path (oracle port, spectra cached as in its perf runs) on this box

--impl reference times that CPU path itself (rank 0 only), same metric/unit/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA7B_LAYER = [("q_proj", 4096, 4096), ("k_proj", 4096, 4096), ("v_proj", 4096, 4096), ("o_proj", 4096, 4096),
                 ("gate_proj", 4096, 11008), ("up_proj", 4096, 11008), ("down_proj", 11008, 4096)]
N_LAYERS = 32
B = 128
METRIC = "bc_linear_decode_tokens_per_s"
UNIT = "tokens/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def config(n_gpus, spec_dtype="float32"):
    """The workload, in the same words for both arms (the driver compares the two `config` objects): per-arm facts (how many
    weight sets rotate on the GPU, which dtype the CPU arm computes in) live in the arms' own top-level keys."""
    return {"workload": "llama2-7b-decode-b128", "projections_per_step": len(LLAMA7B_LAYER) * N_LAYERS, "layers": N_LAYERS,
            "block_size": B, "tokens_per_step_per_gpu": 1, "spectrum_dtype": spec_dtype,
            "weights": "random-init c = randn*0.01 (reference init), distinct per projection (224 matrices)",
            "l2_policy": "inputs larger than L2 (%.0f MB of spectra per step, all distinct)" % (algorithmic_bytes(spec_dtype)[1] / 1e6),
            "parallelism": "dp%d" % n_gpus}


def algorithmic_bytes(spec_dtype="float32", n=1):
    """N*in*s_x + p*q*B*s_w + N*out*s_y per projection; returns (total, spectrum part) per step."""
    sw = 4 if spec_dtype == "float32" else 2
    tot = spec = 0
    for _, inf, outf in LLAMA7B_LAYER:
        spec += inf * outf // B * sw
        tot += n * inf * 2 + inf * outf // B * sw + n * outf * 2
    return tot * N_LAYERS, spec * N_LAYERS


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs: the PCI device's numa_node and the node's cpulist), BEFORE
    any pinned host memory is allocated: first-touch then places the staging buffers on that node.  With one process per GPU and
    unbound processes, 8 ranks x 5 MB per 57 us step cross the socket interconnect and the host-buffer (e2e) figure stops scaling
    (round 1: 0.33 efficiency at 8 GPUs).  Returns a short description for the JSON line; never fails the bench."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"gpu_pci": bdf, "numa_node": node, "bound": False, "why": "platform reports no NUMA affinity"}
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) or allowed
        os.sched_setaffinity(0, use)
        return {"gpu_pci": bdf, "numa_node": node, "bound": bool(cpus & allowed), "cpus": len(use)}
    except Exception as e:                                          # noqa: BLE001
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([s.strip() for s in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the reference's own CPU path (torch.fft.rfft -> einsum -> irfft), restated in
# oracle/bc_oracle.py (bit-identical to the reference on the golden vectors), spectra cached like its perf runs.
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, budget_s=None, cached=True):
    """One decode token through all 224 projections on the host cores, the way the reference's perf runs do it: its own
    BlockCirculantLinear (src/patch_llama_fft.py:142-275) with enable_cfft_cache (src/bench_all_in_one.py:164-254), fp32, batch 1.
    The modules are the UNMODIFIED reference imported from oracle/_ref/ (kind "reference"; staged by oracle/make_ref.py); without it
    the oracle port of the same three torch calls is timed (kind "port").  All 32 layers, distinct weights (207 MB of c + the cached
    complex spectra).  Stops once budget_s seconds of timed work are spent (at least 2 steps).
    Returns (s per token, threads, steps run, kind)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    from oracle import ref_loader
    kind = "reference" if ref_loader.available() else "port"
    mods = []
    if kind == "reference":
        plf, bai = ref_loader.load("patch_llama_fft"), ref_loader.load("bench_all_in_one")
        holder = torch.nn.ModuleList()
        for _ in range(N_LAYERS):
            for _, inf, outf in LLAMA7B_LAYER:
                m = plf.BlockCirculantLinear(inf, outf, block_size=B, bias=False)
                with torch.no_grad():
                    m.c.copy_(torch.randn(outf // B, inf // B, B, generator=g) * 0.01)
                holder.append(m)
        holder.eval()
        if cached:
            bai.enable_cfft_cache(holder)                     # else: the forward as shipped, rfft(c) on every call (src/patch_llama_fft.py:256-257)
        mods = [(m, torch.randn(1, 1, m.in_features, generator=g)) for m in holder]
        call = lambda m, x: m(x)                                                        # noqa: E731
    else:
        from oracle import bc_oracle as O
        for _ in range(N_LAYERS):
            for _, inf, outf in LLAMA7B_LAYER:
                c = torch.randn(outf // B, inf // B, B, generator=g) * 0.01
                mods.append((O.CachedSpectrum(c), torch.randn(1, 1, inf, generator=g)))  # fp32 on CPU, the reference's CPU default
        call = O.bc_forward_cached_torch

    def one_step():
        with torch.inference_mode():
            for m, x in mods:
                call(m, x)
    for _ in range(warmup):
        one_step()
    done = 0
    t0 = time.perf_counter()
    while done < steps:
        one_step()
        done += 1
        if budget_s is not None and done >= 2 and time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return dt, torch.get_num_threads(), done, kind          # seconds per full 32-layer token


CPU_SAMPLE = ("%d timed passes over all 32 layers (224 projections, distinct weights), batch 1, fp32 activations, cached rfft(c) "
              "(enable_cfft_cache, the reference's perf configuration); %s")
CPU_KIND_TEXT = {"reference": "the UNMODIFIED reference modules from oracle/_ref (src/patch_llama_fft.py + src/bench_all_in_one.py)",
                 "port": "oracle/_ref not staged: the oracle port of the reference's CPU path (oracle/bc_oracle.py, torch.fft + einsum)"}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on this box's host cores, all the threads torch will use.
    A step is one full token (all 224 projections); the run stops after ~40 s of timed work however large --steps is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sec, cores, done, kind = cpu_reference_steps(args.steps, min(max(args.warmup, 1), 3), budget_s=40.0)
    val = 1.0 / sec
    sample = CPU_SAMPLE % (done, CPU_KIND_TEXT[kind])
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "decode_us_per_layer": sec * 1e6 / N_LAYERS, "gpu_launches": 0, "activation_dtype": "float32"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import llama_fft_b200 as bc
    from llama_fft_b200 import _cabi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the BC-linear path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    # stdout carries exactly ONE line, the JSON: anything a library writes to fd 1 meanwhile (NCCL prints its version banner there when
    # NCCL_DEBUG is set) goes to stderr instead.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()

    sdt = torch.float32 if args.spectrum == "float32" else torch.bfloat16
    shapes = [(outf // B, inf // B) for _ in range(N_LAYERS) for (_, inf, outf) in LLAMA7B_LAYER]
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    def make_set():
        specs = []
        for (p, q) in shapes:
            c = torch.randn(p, q, B, device=dev, generator=g) * 0.01
            spec = torch.empty(p, q, B, device=dev, dtype=sdt)
            _cabi.check(_cabi.lib().bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.DTYPE[sdt],
                                                            _cabi.LAYOUT_STREAM, _cabi.stream_ptr(dev)), "bcl_precompute_spectrum")
            specs.append(spec)
        grp = bc.GroupedBCForward.from_packed(shapes, specs, 1, torch.float16, dev, num_ctas=args.ctas_per_sm * torch.cuda.get_device_properties(dev).multi_processor_count)
        grp.x_all.copy_(torch.randn(grp.x_all.numel(), device=dev, generator=g).to(torch.float16))
        return grp

    groups = [make_set() for _ in range(args.weight_sets)]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, one grouped launch each ---------------------------------------------------
    for i in range(args.warmup):
        groups[i % len(groups)].run()
    barrier()
    # The timed region is REPEATS back-to-back blocks of exactly --steps steps, each block bracketed by its own CUDA-event pair on the
    # launching stream, the whole region by barrier + synchronize; per block the max over ranks, then the MEDIAN block (a 20-step
    # block is 0.8 ms: one event pair alone reads +-4 % from run to run).
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    blocks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.repeats)]
    with ClockSampler(local) as clk:
        for (b0, b1) in blocks:
            b0.record(stream)
            for i in range(args.steps):
                groups[i % len(groups)].run()
            b1.record(stream)
        barrier()
    t = torch.tensor([b0.elapsed_time(b1) for (b0, b1) in blocks], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    block_ms = sorted(t.tolist())
    ms_step = block_ms[len(block_ms) // 2] / args.steps
    launches = args.steps * args.repeats

    # ---- end to end through the host-buffer entry point ------------------------------------------------------------
    # Every step: H2D of that step's x from pinned memory, one grouped launch, D2H of its y to pinned memory, through the
    # C ABI's pipelined host call (bcl_forward_grouped_host_pipelined): three streams, the weight sets double as staging
    # slots, so copies of neighbouring steps overlap the kernel.  The host consumes (waits for) a slot's previous result
    # before resubmitting it, so at most len(groups) steps are in flight and every result reaches the host inside the region.
    n_slots = 4                                               # staging slots (own x/y buffers + plan); slot j streams weight set j % sets (3: 75 us, 6-8: 57-61 us)
    slots = list(groups)
    while len(slots) < n_slots:
        src = groups[len(slots) % len(groups)]
        slots.append(bc.GroupedBCForward.from_packed(shapes, src.specs, 1, torch.float16, dev, num_ctas=src.num_ctas))
    hx = [torch.randn(grp.x_all.numel()).to(torch.float16).pin_memory() for grp in slots]
    hy = [torch.empty(grp.y_all.numel(), dtype=torch.float16).pin_memory() for grp in slots]
    pipe = bc.HostPipe(dev)
    for i in range(min(args.warmup, 8)):
        k = i % n_slots
        slots[k].wait_host()
        slots[k].submit_host(hx[k], hy[k], pipe)
    pipe.synchronize()
    barrier()
    e2e_steps = max(10, min(args.steps, 2000))
    ev0.record(stream)
    for s in (pipe.h2d, pipe.compute, pipe.d2h):
        s.wait_event(ev0)
    for i in range(e2e_steps):
        k = i % n_slots
        slots[k].wait_host()
        slots[k].submit_host(hx[k], hy[k], pipe)
    for s in (pipe.h2d, pipe.compute, pipe.d2h):
        fin = torch.cuda.Event()
        fin.record(s)
        stream.wait_event(fin)
    ev1.record(stream)
    barrier()
    t2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = t2.item() / e2e_steps
    # the same steps with ONE stream per staging slot (bcl_forward_grouped_host_async: three API calls per step, no events): the
    # copy engines and the SMs overlap neighbouring slots by themselves; the better of the two forms is the e2e number
    slot_streams = [torch.cuda.Stream(dev) for _ in range(n_slots)]
    for i in range(min(args.warmup, 8)):
        k = i % n_slots
        slot_streams[k].synchronize()
        slots[k].submit_host_stream(hx[k], hy[k], slot_streams[k])
    for s in slot_streams:
        s.synchronize()
    barrier()
    ev0.record(stream)
    for s in slot_streams:
        s.wait_event(ev0)
    for i in range(e2e_steps):
        k = i % n_slots
        slot_streams[k].synchronize()                         # the host has this slot's previous result
        slots[k].submit_host_stream(hx[k], hy[k], slot_streams[k])
    for s in slot_streams:
        fin = torch.cuda.Event()
        fin.record(s)
        stream.wait_event(fin)
    ev1.record(stream)
    barrier()
    t3 = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
    e2e_ms_three_streams = e2e_ms
    e2e_ms_slot_streams = t3.item() / e2e_steps
    e2e_ms = min(e2e_ms_three_streams, e2e_ms_slot_streams)
    # PCIe copies alone (what bounds a pipelined step): 20 back-to-back copies each way
    ev0.record(stream)
    for _ in range(20):
        slots[0].x_all.copy_(hx[0], non_blocking=True)
    ev1.record(stream)
    torch.cuda.synchronize()
    h2d_us = ev0.elapsed_time(ev1) / 20 * 1e3
    ev0.record(stream)
    for _ in range(20):
        hy[0].copy_(slots[0].y_all, non_blocking=True)
    ev1.record(stream)
    torch.cuda.synchronize()
    d2h_us = ev0.elapsed_time(ev1) / 20 * 1e3
    # the same step without overlap (one stream, synchronise per step), for the record
    for i in range(3):
        groups[i % len(groups)].run_host(hx[i % len(groups)], hy[i % len(groups)])
    barrier()
    ev0.record(stream)
    for i in range(50):
        k = i % len(groups)
        groups[k].run_host(hx[k], hy[k])
    ev1.record(stream)
    barrier()
    e2e_serial_ms = ev0.elapsed_time(ev1) / 50

    # ---- BASELINE.json configs[4]: Llama-2-70B shapes, output-block-row sharded over the N GPUs of this run (all ranks take part) -----
    sharded = None
    if world > 1 and not args.no_detail:
        try:
            sharded = sharded_70b_detail(bc, dev, world, rank)
        except Exception as e:                      # noqa: BLE001 -- evidence only; never hides the headline line
            sharded = {"error": "%s: %s" % (type(e).__name__, e)}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- extra evidence at N=1 (not part of the timed value): other launch granularities, bf16 spectra --------------
    detail = {}
    if sharded is not None:
        detail["sharded_70b"] = sharded
    if world == 1 and not args.no_detail:
        detail = launch_granularity_detail(bc, _cabi, groups[0], shapes, dev, args)
        try:
            detail["decode_independent_launches"] = independent_launch_detail(bc, groups, dev, shapes, algorithmic_bytes(args.spectrum)[0])
        except Exception as e:                                           # noqa: BLE001
            detail["decode_independent_launches"] = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            detail["decode_bf16_spectra"] = bf16_decode_detail(bc, _cabi, dev, shapes, args)
        except Exception as e:                      # noqa: BLE001
            detail["decode_bf16_spectra"] = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            detail["prefill"] = prefill_detail(bc, _cabi, dev)
        except Exception as e:                      # noqa: BLE001 -- evidence only; never hides the headline line
            detail["prefill"] = {"error": "%s: %s" % (type(e).__name__, e)}
        try:
            detail["reference_cuda"] = reference_cuda_detail(dev)
        except Exception as e:                      # noqa: BLE001
            detail["reference_cuda"] = {"error": "%s: %s" % (type(e).__name__, e)}

    peak, peak_src = peaks()
    alg, spec_bytes = algorithmic_bytes(args.spectrum)
    achieved = alg / (ms_step * 1e-3) / 1e9
    # DRAM bytes per launch of the hot kernel cannot be read without a profiler: the figure is the one ncu --set full capture of
    # this command committed under profiles/ (newest round first), and `traffic_source` says which file it came from.
    traffic, traffic_src = None, None
    for name in ("r2_traffic.json", "traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("stream_ws_kernel_%s" % args.spectrum)
            if traffic is not None:
                traffic_src = "profiles/" + name + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
                break
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sec, cores, done, kind = cpu_reference_steps(50, 1, budget_s=15.0)
        cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": CPU_SAMPLE % (done, CPU_KIND_TEXT[kind])}
        if kind == "reference":                               # also the forward as shipped (no spectrum cache), a few passes
            sec_u, _, done_u, _ = cpu_reference_steps(8, 1, budget_s=6.0, cached=False)
            cpu["value_as_shipped_uncached"] = 1.0 / sec_u
            cpu["sample_as_shipped"] = "%d timed passes, rfft(c) recomputed by every forward (src/patch_llama_fft.py:256-257)" % done_u
    line = {"metric": METRIC, "value": world * 1.0 / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(world, args.spectrum),
            "decode_us_per_layer": ms_step * 1e3 / N_LAYERS, "activation_dtype": "float16", "weight_sets_alternating": args.weight_sets,
            "timing": {"repeats": args.repeats, "steps_per_block": args.steps, "block_ms_sorted": block_ms, "reported": "median block / steps"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "bcl::ws::stream_ws_kernel<128,%s,3> (warp-specialised TMA/mbarrier pipeline, 1 CTA/SM)" % ("float" if args.spectrum == "float32" else "bf16"),
                         "algorithmic_bytes_per_launch": alg, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": world * 1.0 / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": groups[0].x_all.numel() * 2, "d2h_bytes_per_step": groups[0].y_all.numel() * 2,
                    "how": "pinned x -> H2D -> grouped launch -> D2H -> pinned y every step, %d staging slots in flight, the better of "
                           "bcl_forward_grouped_host_async with one stream per slot (%.4f ms/step) and bcl_forward_grouped_host_pipelined on three "
                           "shared streams (%.4f ms/step); copies alone: H2D %.1f us, D2H %.1f us per step; unpipelined (bcl_forward_grouped_host, "
                           "sync per step): %.4f ms/step" % (n_slots, e2e_ms_slot_streams, e2e_ms_three_streams, h2d_us, d2h_us, e2e_serial_ms)},
            "gpu_launches": launches, "clocks": clk.summary(), "host_numa_binding": numa, "detail": detail}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def launch_granularity_detail(bc, _cabi, grp_all, shapes, dev, args):
    """How the same 224 projections fare at the launch granularities a real decoder allows."""
    import torch
    out = {}
    stream = torch.cuda.current_stream(dev)
    specs, xs = grp_all.specs, grp_all.xs

    def time_graph(fn, reps=20):
        s = torch.cuda.Stream(dev)
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                fn()
            for _ in range(3):
                graph.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(reps):
                graph.replay()
            e1.record(s)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    L = _cabi.lib()
    per_layer = [bc.GroupedBCForward(xs[7 * l:7 * l + 7], specs[7 * l:7 * l + 7]) for l in range(N_LAYERS)]
    out["per_layer_grouped_graph_ms_per_step"] = time_graph(lambda: [gl.run() for gl in per_layer])
    out["per_layer_grouped_launches_per_step"] = N_LAYERS
    ys = [torch.empty(1, p * B, device=dev, dtype=torch.float16) for (p, q) in shapes]

    def per_module():
        st = torch.cuda.current_stream(dev).cuda_stream
        for x, s, y, (p, q) in zip(xs, specs, ys, shapes):
            L.bcl_forward(x.data_ptr(), _cabi.F16, 1, q * B, s.data_ptr(), _cabi.DTYPE[s.dtype], _cabi.LAYOUT_STREAM, p, q, B,
                          None, 0, y.data_ptr(), _cabi.F16, p * B, None, 0, st)
    out["per_module_graph_ms_per_step"] = time_graph(per_module)
    out["per_module_launches_per_step"] = len(shapes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_module(); torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(5):
        per_module()
    e1.record(stream); torch.cuda.synchronize()
    out["per_module_eager_ms_per_step"] = e0.elapsed_time(e1) / 5
    # the empty-kernel launch floor under the same graph replay: what 224 / 32 launches cost when they do nothing
    def noops(n, pdl):
        st = torch.cuda.current_stream(dev).cuda_stream
        for _ in range(n):
            L.bcl_launch_noop(pdl, st)
    floor_pdl = time_graph(lambda: noops(len(shapes), 1)) / len(shapes) * 1e3
    floor_plain = time_graph(lambda: noops(len(shapes), 0)) / len(shapes) * 1e3
    out["launch_floor"] = {"what": "CUDA-graph replay of %d empty kernels (1 CTA x 32 threads) back to back on one stream" % len(shapes),
                           "empty_kernel_us_per_launch_pdl": floor_pdl, "empty_kernel_us_per_launch_plain": floor_plain,
                           "per_module_us_per_launch": out["per_module_graph_ms_per_step"] * 1e3 / len(shapes),
                           "per_layer_grouped_us_per_launch": out["per_layer_grouped_graph_ms_per_step"] * 1e3 / N_LAYERS}
    try:
        out["dependent_chain"] = dependent_chain_detail(bc, _cabi, dev, time_graph)
    except Exception as e:                                               # noqa: BLE001
        out["dependent_chain"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return out


def bf16_decode_detail(bc, _cabi, dev, shapes, args, steps=1000):
    """The headline workload with bf16 spectra (half the bytes, parity class 1e-2): same grouped launch, two weight sets."""
    import torch
    g = torch.Generator(device=dev).manual_seed(4321)
    groups = []
    for _ in range(2):
        specs = []
        for (p, q) in shapes:
            c = torch.randn(p, q, B, device=dev, generator=g) * 0.01
            spec = torch.empty(p, q, B, device=dev, dtype=torch.bfloat16)
            _cabi.check(_cabi.lib().bcl_precompute_spectrum(c.data_ptr(), p, q, B, spec.data_ptr(), _cabi.BF16, _cabi.LAYOUT_STREAM,
                                                            _cabi.stream_ptr(dev)), "bcl_precompute_spectrum")
            specs.append(spec)
        grp = bc.GroupedBCForward.from_packed(shapes, specs, 1, torch.float16, dev, num_ctas=args.ctas_per_sm * torch.cuda.get_device_properties(dev).multi_processor_count)
        grp.x_all.copy_(torch.randn(grp.x_all.numel(), device=dev, generator=g).to(torch.float16))
        groups.append(grp)
    for i in range(20):
        groups[i & 1].run()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        groups[i & 1].run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    alg, _ = algorithmic_bytes("bfloat16")
    peak, _ = peaks()
    del groups
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "tokens_per_s": 1.0 / (ms * 1e-3), "algorithmic_bytes_per_launch": alg, "hbm_frac_of_measured_peak": alg / (ms * 1e-3) / 1e9 / peak,
            "note": "row-per-warp kernel (bcl_stream_rw.cu: packed FFMA2 MAC loop, 3 TMA producer warps, 20-warp CTA): with 8 KB rows the split-q kernel of the headline is bound by its per-row hand-offs, not by HBM; DESIGN 3.3b"}


def independent_launch_detail(bc, groups, dev, shapes, alg, steps=2000):
    """Where the headline's last ~10 % go.  Every launch of the headline calls griddepcontrol.wait before it touches activations (the
    previous launch may have produced them), so step k+1 computes only after the SLOWEST CTA of step k has finished.  The bench's
    steps do not depend on each other (two alternating buffer sets); plans encoded with BCL_GROUP_INDEPENDENT=1 skip the wait, an SM
    starts the next step the moment its CTA leaves, and the same kernel runs at the rate its stream sustains.  A measurement switch:
    a decoder's consecutive tokens DO depend on each other, so the headline keeps the wait."""
    import torch
    os.environ["BCL_GROUP_INDEPENDENT"] = "1"
    try:
        ind = [bc.GroupedBCForward.from_packed(shapes, grp.specs, 1, torch.float16, dev, num_ctas=grp.num_ctas) for grp in groups]
    finally:
        os.environ.pop("BCL_GROUP_INDEPENDENT", None)
    for g2, g1 in zip(ind, groups):
        g2.x_all.copy_(g1.x_all)
    for i in range(20):
        ind[i % len(ind)].run()
    torch.cuda.synchronize()
    same = all(bool(torch.equal(g2.y_all, g1.y_all)) for g2, g1 in zip(ind, groups))
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        ind[i % len(ind)].run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    peak, _ = peaks()
    return {"what": "the headline launches without griddepcontrol.wait (BCL_GROUP_INDEPENDENT=1 plans): consecutive steps overlap on the SMs",
            "ms_per_step": ms, "tokens_per_s": 1.0 / (ms * 1e-3), "hbm_frac_of_measured_peak": alg / (ms * 1e-3) / 1e9 / peak,
            "outputs_bit_identical_to_headline": same}


def dependent_chain_detail(bc, _cabi, dev, time_graph):
    """What a real decoder can do: inside a layer the projections depend on each other (q/k/v -> o -> gate/up -> down), so a token
    costs 4 DEPENDENT launches per layer with shared-input fusion (q/k/v and gate/up stacked), 7 without.  CUDA-graph replay of
    32 layers with distinct weights (no L2 reuse), each launch reading the previous launch's output; next to it the dense fp16
    nn.Linear chain of the same shapes on the same GPU (cuBLAS; weight-bandwidth bound at batch 1)."""
    import torch
    torch.manual_seed(0)
    sdt = torch.float32

    def mk(inf, outf):
        return bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=sdt).to(dev).eval()
    layers = []
    for _ in range(N_LAYERS):
        qkv, o, gu, down = mk(4096, 3 * 4096), mk(4096, 4096), mk(4096, 2 * 11008), mk(11008, 4096)
        for m in (qkv, o, gu, down):
            m.build_spectrum()
        layers.append((qkv, o, gu, down))
    x0 = torch.randn(1, 4096, device=dev).half()

    def bc_chain():
        h = x0
        with torch.no_grad():
            for qkv, o, gu, down in layers:
                a = qkv(h)[:, :4096]                      # stand-in for attention: any slice keeps the dependency
                h = o(a.contiguous())
                g = gu(h)
                h = down(g[:, :11008].contiguous())
        return h
    bc_ms = time_graph(bc_chain, reps=10)
    # the same dependency chain on the chain kernel (bcl_forward_chain): q/k/v and o as single-step launches, the MLP as ONE launch
    # (gate|up with the SWIGLU epilogue -> grid barrier -> down, rfft(h) handed over as a spectrum), and -- as the bound of what fusing
    # can give -- the whole layer as one 4-step launch (a real decoder has attention between q/k/v and o)
    from llama_fft_b200 import _cabi as C
    f16 = dict(device=dev, dtype=torch.float16)
    H, I = 4096, 11008
    buf = dict(qkv=torch.empty(1, 3 * H, **f16), a=torch.empty(1, H, **f16), h=torch.empty(1, I, **f16), hf=torch.empty(1, I, device=dev, dtype=torch.float32),
               y=[torch.empty(1, H, **f16) for _ in range(2)])
    three, whole, cur = [], [], x0
    for k, (qkv, o, gu, down) in enumerate(layers):
        sq, so, sg, sd = (m.build_spectrum() for m in (qkv, o, gu, down))
        y = buf["y"][k & 1]
        mlp = [[(buf["a"], sg, buf["h"], C.EPILOGUE_SWIGLU, None, buf["hf"])], [(buf["hf"], sd, y, C.EPILOGUE_LINEAR, None, None, True)]]
        three += [bc.BCChain([[(cur, sq, buf["qkv"], 0)]]), bc.BCChain([[(buf["qkv"][:, :H], so, buf["a"], 0)]]), bc.BCChain(mlp)]
        whole.append(bc.BCChain([[(cur, sq, buf["qkv"], 0)], [(buf["qkv"][:, :H], so, buf["a"], 0)]] + mlp))
        cur = y
    three_ms = time_graph(lambda: [c.run() for c in three], reps=10)
    whole_ms = time_graph(lambda: [c.run() for c in whole], reps=10)
    layer_bytes = algorithmic_bytes("float32")[0] / N_LAYERS
    peak, _ = peaks()
    res = {"what": "32 layers of distinct weights, 1 token, fp32 spectra, CUDA-graph replay, every launch reads the previous launch's output "
                   "(q/k/v and gate/up stacked: shared-input fusion).  bc_us_per_layer = 3 launches per layer on the chain kernel: q/k/v, o, "
                   "fused MLP (gate|up + SiLU*mul + down in one launch); per_module = 4 launches of the per-module kernel (round 1's figure); "
                   "whole_layer = all four steps in one launch (synthetic: attention sits between q/k/v and o in a real decoder)",
           "bc_ms_per_token": three_ms, "bc_us_per_layer": three_ms * 1e3 / N_LAYERS,
           "bc_us_per_layer_per_module_kernel_4_launches": bc_ms * 1e3 / N_LAYERS,
           "bc_us_per_layer_whole_layer_one_launch": whole_ms * 1e3 / N_LAYERS,
           "hbm_frac_of_measured_peak": layer_bytes / (three_ms * 1e-3 / N_LAYERS) / 1e9 / peak,
           "algorithmic_bytes_per_layer": layer_bytes}
    # end to end for ONE token through that chain, moving what a real token moves: 8 KB of hidden state in (pinned host -> device),
    # 96 launches (one CUDA-graph replay), 8 KB of hidden state out, stream synchronised -- every token
    try:
        hx = torch.randn(1, H).half().pin_memory()
        hy = torch.empty(1, H, dtype=torch.float16).pin_memory()
        s2 = torch.cuda.Stream(dev)
        with torch.cuda.stream(s2):
            for c in three:
                c.run()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s2):
                for c in three:
                    c.run()

            def token():
                x0.copy_(hx, non_blocking=True)
                graph.replay()
                hy.copy_(cur, non_blocking=True)
                s2.synchronize()
            for _ in range(5):
                token()
            t0 = time.perf_counter()
            for _ in range(200):
                token()
            tok_s = 200 / (time.perf_counter() - t0)
        res["e2e_token_path"] = {"tokens_per_s": tok_s, "us_per_token": 1e6 / tok_s, "h2d_bytes_per_token": H * 2, "d2h_bytes_per_token": H * 2,
                                 "what": "per token: pinned hidden state -> device, one graph replay of the 96-launch chain, hidden state -> pinned host, synchronise (wall clock)"}
    except Exception as e:                                               # noqa: BLE001
        res["e2e_token_path"] = {"error": "%s: %s" % (type(e).__name__, e)}
    bc_ms = three_ms
    del three, whole
    del layers
    torch.cuda.empty_cache()
    dense = []
    for _ in range(N_LAYERS):
        dense.append(tuple(torch.nn.Linear(i, o_, bias=False, device=dev, dtype=torch.float16) for i, o_ in
                           ((4096, 3 * 4096), (4096, 4096), (4096, 2 * 11008), (11008, 4096))))

    def dense_chain():
        h = x0
        with torch.no_grad():
            for qkv, o, gu, down in dense:
                a = qkv(h)[:, :4096]
                h = o(a.contiguous())
                g = gu(h)
                h = down(g[:, :11008].contiguous())
        return h
    d_ms = time_graph(dense_chain, reps=10)
    res.update({"dense_fp16_ms_per_token": d_ms, "dense_fp16_us_per_layer": d_ms * 1e3 / N_LAYERS, "speedup_vs_dense_fp16": d_ms / bc_ms})
    del dense
    torch.cuda.empty_cache()
    return res


PREFILL_TOKENS = 8 * 2048          # BASELINE.json configs[2]: 2048-token prefill x batch 8

SHAPES_70B = [("q/o", 8192, 8192), ("k/v", 8192, 1024), ("gate/up", 8192, 28672), ("down", 28672, 8192)]


def sharded_70b_detail(bc, dev, world, rank, B70=256, reps=10):
    """BASELINE.json configs[4]: Llama-2-70B projection shapes, B = 256, bf16 spectra, fp16 activations, output block-rows sharded over
    the `world` GPUs of this run (llama-fft_b200/sharded.py), at 1 token (decode) and 2048 tokens (prefill).  Per shape: the unsharded
    layer on one GPU, the rank-local slice alone, and the complete sharded forward with each exchange -- NCCL all-gather, push (slice
    written in place, one bcl_push_columns launch over NVLink), p2p (stores to every peer fused into the producing kernels) -- with
    the result checked against the unsharded layer on the same GPU.  CUDA events, max over ranks.  exchange_GBps = bytes a rank
    RECEIVES, (G-1)/G * N * out * 2, over (total - local compute)."""
    import torch
    import torch.distributed as dist

    def ev_time(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    rows = []
    for name, inf, outf in SHAPES_70B:
        torch.manual_seed(0)                                        # same weights on every rank
        full = bc.BlockCirculantLinear(inf, outf, block_size=B70, bias=False, spectrum_dtype=torch.bfloat16).to(dev).eval()
        shards = {}
        for mode in ("nccl", "push", "p2p"):
            try:
                shards[mode] = bc.shard_bc_linear(full, exchange=mode)
            except Exception as e:                                  # noqa: BLE001
                shards[mode] = "%s: %s" % (type(e).__name__, e)
        for ntok in (1, 2048):
            x = torch.randn(ntok, inf, device=dev, generator=torch.Generator(device=dev).manual_seed(1)).half()
            row = {"shape": "%dx%d" % (inf, outf), "name": name, "tokens": ntok, "world": world,
                   "rows_per_rank": [j1 - j0 for j0, j1 in bc.shard_rows(outf // B70, world)]}
            with torch.no_grad():
                y_full = full(x)
                row["unsharded_us"] = ev_time(lambda: full(x))
                loc = shards["nccl"].local if not isinstance(shards["nccl"], str) else None
                row["local_compute_us"] = ev_time((lambda: loc(x)) if loc is not None else (lambda: None))
                recv_bytes = (world - 1) / world * ntok * outf * 2
                for mode, sh in shards.items():
                    if isinstance(sh, str):
                        row[mode + "_error"] = sh[:200]
                        continue
                    try:
                        y = sh(x)
                        torch.cuda.synchronize()
                        err = (y.float() - y_full.float()).abs().max().item() / max(y_full.float().abs().max().item(), 1e-30)
                        t = torch.tensor([err], device=dev, dtype=torch.float64)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        row[mode + "_rel_err_vs_unsharded"] = t.item()
                        row[mode + "_total_us"] = ev_time(lambda: sh(x))
                        ex_us = max(row[mode + "_total_us"] - row["local_compute_us"], 1e-3)
                        row[mode + "_exchange_us"] = ex_us
                        row[mode + "_exchange_GBps"] = recv_bytes / (ex_us * 1e-6) / 1e9
                    except Exception as e:                          # noqa: BLE001
                        row[mode + "_error"] = ("%s: %s" % (type(e).__name__, e))[:200]
            rows.append({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()})
        del full, shards
        torch.cuda.empty_cache()
    return {"what": "Llama-2-70B projection shapes, B = 256, bf16 spectra, fp16 activations, output block-rows sharded over %d GPUs; us, max over ranks" % world,
            "rows": rows}


def reference_cuda_detail(dev):
    """The kernel-level bar of SURVEY 8(d) / BASELINE.md section 4: the UNMODIFIED reference module (oracle/_ref; torch.fft = cuFFT,
    einsum = cuBLAS complex bmm) on the same B200 and the same shapes, as shipped (rfft(c) on every call, src/patch_llama_fft.py:256-257)
    and with enable_cfft_cache (src/bench_all_in_one.py:164-254), at decode (1 token) and at the prefill size (16384 tokens).
    fp16 activations, fp32 parameters -- the reference's Llama configuration (commands.txt:22)."""
    import torch
    from oracle import ref_loader
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref not staged (python oracle/make_ref.py in the build container)"}
    plf, bai = ref_loader.load("patch_llama_fft"), ref_loader.load("bench_all_in_one")
    stream = torch.cuda.current_stream(dev)
    out = {"what": "reference BlockCirculantLinear on cuda (torch.fft + einsum), us per call; layer = 4 x 4096->4096 + 2 x 4096->11008 + 11008->4096"}
    torch.manual_seed(0)
    for n, reps in ((1, 50), (PREFILL_TOKENS, 3)):
        row = {}
        for cached in (0, 1):
            per = {}
            for inf, outf in ((4096, 4096), (4096, 11008), (11008, 4096)):
                m = plf.BlockCirculantLinear(inf, outf, block_size=B, bias=False).to(dev).eval()
                holder = torch.nn.ModuleList([m])
                if cached:
                    bai.enable_cfft_cache(holder)
                x = torch.randn(n, inf, device=dev).half()
                with torch.inference_mode():
                    for _ in range(3):
                        m(x)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(reps):
                        m(x)
                    e1.record(stream)
                    torch.cuda.synchronize()
                per["%dx%d" % (inf, outf)] = e0.elapsed_time(e1) / reps * 1e3
                del m, holder, x
                torch.cuda.empty_cache()
            per["layer_us"] = 4 * per["4096x4096"] + 2 * per["4096x11008"] + per["11008x4096"]
            row["cached_rfft_c" if cached else "as_shipped"] = per
        out["decode_1_token" if n == 1 else "prefill_%d_tokens" % n] = row
    return out


def prefill_algorithmic_bytes(n=PREFILL_TOKENS):
    """SURVEY 8(d) with fp16 activations and bf16 spectra, one decoder layer (7 projections)."""
    return sum(n * inf * 2 + inf * outf // B * 2 + n * outf * 2 for _, inf, outf in LLAMA7B_LAYER)


def prefill_algorithmic_flops(n=PREFILL_TOKENS):
    """4 N p q (B-1) per projection (complex MACs of the per-frequency contraction; DFT flops excluded)."""
    return sum(4 * n * (outf // B) * (inf // B) * (B - 1) for _, inf, outf in LLAMA7B_LAYER)


def time_prefill_layer(bc, _cabi, dev, reps, n=PREFILL_TOKENS, through_host=False, spectrum_dtype=None):
    """One decoder layer's 7 BC projections (bf16 spectra -> tcgen05 prefill path) on n tokens of synthetic fp16
    activations; inputs rotate over two buffer sets (each larger than L2).  Returns (us per layer, per-shape us, status)."""
    import torch
    from llama_fft_b200.patch_llama_fft import bc_forward, prefill_status
    torch.manual_seed(0)
    sdt = torch.bfloat16 if spectrum_dtype is None else spectrum_dtype
    layers = {}
    for name, inf, outf in LLAMA7B_LAYER:
        if (inf, outf) not in layers:
            m = bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=sdt).to(dev).eval()
            layers[(inf, outf)] = (m, m.build_spectrum(layout=_cabi.LAYOUT_GEMM))
    xs = {d: [torch.randn(n, d, device=dev).half() for _ in range(2)] for d in (4096, 11008)}
    ys = {d: [torch.empty(n, d, device=dev, dtype=torch.float16) for _ in range(2)] for d in (4096, 11008)}
    hx = hy = None
    if through_host:
        hx = {d: torch.randn(n, d).half().pin_memory() for d in (4096, 11008)}
        hy = {d: torch.empty(n, d, dtype=torch.float16).pin_memory() for d in (4096, 11008)}
    stream = torch.cuda.current_stream(dev)

    def one(i, inf, outf):
        m, spec = layers[(inf, outf)]
        x, y = xs[inf][i & 1], ys[outf][i & 1]
        if through_host:
            x.copy_(hx[inf], non_blocking=True)
        bc_forward(x, spec, None, outf // B, inf // B, B, out=y, layout=_cabi.LAYOUT_GEMM, spectrum_dtype=sdt)
        if through_host:
            hy[outf].copy_(y, non_blocking=True)
            stream.synchronize()

    per_shape = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for (inf, outf) in layers:
        for i in range(3):
            one(i, inf, outf)
        torch.cuda.synchronize()
        e0.record(stream)
        for i in range(reps):
            one(i, inf, outf)
        e1.record(stream)
        torch.cuda.synchronize()
        per_shape["%dx%d" % (inf, outf)] = e0.elapsed_time(e1) / reps * 1e3
    e0.record(stream)
    for i in range(reps):
        for _, inf, outf in LLAMA7B_LAYER:
            one(i, inf, outf)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, per_shape, prefill_status(dev)


def time_prefill_layer_shared_input(bc, _cabi, dev, reps, n=PREFILL_TOKENS):
    """Same layer, same tokens, with q/k/v and gate/up fused into one BC forward each (llama-fft_b200/shared_input.py):
    4 launches-chains instead of 7, one forward-DFT pass of x per group.  The algorithmic bytes of the fused layer are smaller
    (x is read once per group), so the roofline fraction is quoted against its own figure."""
    import torch
    from llama_fft_b200.patch_llama_fft import prefill_status
    torch.manual_seed(0)

    def mk(inf, outf):
        return bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=torch.bfloat16).to(dev).eval()
    qkv = bc.SharedInputGroup([mk(4096, 4096) for _ in range(3)])
    gu = bc.SharedInputGroup([mk(4096, 11008) for _ in range(2)])
    o, down = mk(4096, 4096), mk(11008, 4096)
    xs = {d: [torch.randn(n, d, device=dev).half() for _ in range(2)] for d in (4096, 11008)}

    def layer(i):
        with torch.no_grad():
            qkv.forward_all(xs[4096][i & 1])
            o(xs[4096][i & 1])
            gu.forward_all(xs[4096][i & 1])
            down(xs[11008][i & 1])
    for i in range(3):
        layer(i)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        layer(i)
    e1.record(stream)
    torch.cuda.synchronize()
    alg = sum(n * inf * 2 + inf * outf // B * 2 + n * outf * 2 for inf, outf in ((4096, 3 * 4096), (4096, 4096), (4096, 2 * 11008), (11008, 4096)))
    return e0.elapsed_time(e1) / reps * 1e3, alg, prefill_status(dev)


def time_prefill_mlp(bc, _cabi, dev, reps, n=PREFILL_TOKENS, spectrum_dtype=None):
    """The MLP of one layer at prefill: what HF's LlamaMLP.forward runs over the three BC modules (gate, up, act_fn, multiply, down),
    the same with gate | up stacked (shared-input fusion), and the whole-MLP fusion (bcl_forward_swiglu: SWIGLU epilogue inside the
    tensor-core inverse stage, then down_proj)."""
    import torch
    import torch.nn.functional as F
    from llama_fft_b200.patch_llama_fft import prefill_status
    torch.manual_seed(0)
    sdt = spectrum_dtype or torch.bfloat16

    def mk(inf, outf):
        return bc.BlockCirculantLinear(inf, outf, block_size=B, bias=False, spectrum_dtype=sdt).to(dev).eval()
    gate, up, down = mk(4096, 11008), mk(4096, 11008), mk(11008, 4096)
    fused = bc.FusedBCMLP(gate, up, down)
    xs = [torch.randn(n, 4096, device=dev).half() for _ in range(2)]
    stream = torch.cuda.current_stream(dev)

    def timed(fn):
        with torch.no_grad():
            for i in range(2):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(reps):
                fn(i)
            e1.record(stream)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    t_modules = timed(lambda i: down(F.silu(gate(xs[i & 1])) * up(xs[i & 1])))

    def stacked(i):
        gu = fused.group.forward_all(xs[i & 1])
        return down(F.silu(gu[:, :11008]) * gu[:, 11008:])
    t_stacked = timed(stacked)
    t_fused = timed(lambda i: fused.forward_prefill(xs[i & 1]))
    with torch.no_grad():
        a, b = fused.forward_prefill(xs[0]).float(), stacked(0).float()
    rel = ((a - b).abs().max() / b.abs().max()).item()
    # bytes the fused MLP has to move: x in, h out and in again (down_proj is a separate forward), y out, + the spectra
    alg = n * 4096 * 2 + 2 * n * 11008 * 2 + n * 4096 * 2 + 3 * 4096 * 11008 // B * 2
    return {"what": "one layer's MLP, %d tokens, fp16, %s spectra: down(silu(gate(x)) * up(x))" % (n, "bf16" if sdt == torch.bfloat16 else "fp32 (split hi/lo)"),
            "three_modules_plus_torch_silu_mul_us": t_modules, "stacked_gate_up_plus_torch_silu_mul_us": t_stacked,
            "fused_swiglu_stage_us": t_fused, "speedup_vs_modules": t_modules / t_fused, "max_rel_diff_vs_stacked": rel,
            "algorithmic_bytes_fused": alg, "pipeline_status": prefill_status(dev)}


def prefill_detail(bc, _cabi, dev, reps=5):
    import torch
    us, per_shape, status = time_prefill_layer(bc, _cabi, dev, reps)
    peak, _ = peaks()
    alg, fl = prefill_algorithmic_bytes(), prefill_algorithmic_flops()
    tpeak = 1685.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tpeak = float(json.load(f)["bf16_tflops"])
    except Exception:
        pass
    out = {"workload": "llama2-7b-prefill-b128: 7 projections of one layer, %d tokens, fp16 activations, bf16 spectra (tcgen05 path)" % PREFILL_TOKENS,
           "us_per_layer": us, "tokens_per_s": PREFILL_TOKENS / (us * 1e-6), "us_per_projection_shape": per_shape,
           "algorithmic_bytes_per_layer": alg, "hbm_frac_of_measured_peak": alg / (us * 1e-6) / 1e9 / peak,
           "algorithmic_tflops": fl / (us * 1e-6) / 1e12, "tensor_frac_of_measured_bf16_peak": fl / (us * 1e-6) / 1e12 / tpeak,
           "pipeline_status": status}
    torch.cuda.empty_cache()
    try:
        us2, alg2, status2 = time_prefill_layer_shared_input(bc, _cabi, dev, reps)
        out["shared_input_fusion"] = {"what": "q/k/v and gate/up each as one BC forward over stacked block-rows (bit-identical outputs)",
                                      "us_per_layer": us2, "tokens_per_s": PREFILL_TOKENS / (us2 * 1e-6), "algorithmic_bytes_per_layer": alg2,
                                      "hbm_frac_of_measured_peak": alg2 / (us2 * 1e-6) / 1e9 / peak, "pipeline_status": status2}
    except Exception as e:                                               # noqa: BLE001
        out["shared_input_fusion"] = {"error": "%s: %s" % (type(e).__name__, e)}
    torch.cuda.empty_cache()
    try:                                                                 # the default spectrum_dtype (fp32): split (hi, lo) operands on tcgen05, 1e-4 class
        us3, per3, status3 = time_prefill_layer(bc, _cabi, dev, max(2, reps // 2), spectrum_dtype=torch.float32)
        alg3 = sum(PREFILL_TOKENS * inf * 2 + inf * outf // B * 4 + PREFILL_TOKENS * outf * 2 for _, inf, outf in LLAMA7B_LAYER)
        out["fp32_spectra"] = {"what": "same layer with fp32 spectra (the default spectrum_dtype): split (hi, lo) bf16 operands, fp32 accumulation, 1e-4 parity class",
                               "us_per_layer": us3, "tokens_per_s": PREFILL_TOKENS / (us3 * 1e-6), "us_per_projection_shape": per3,
                               "algorithmic_bytes_per_layer": alg3, "hbm_frac_of_measured_peak": alg3 / (us3 * 1e-6) / 1e9 / peak, "pipeline_status": status3}
    except Exception as e:                                               # noqa: BLE001
        out["fp32_spectra"] = {"error": "%s: %s" % (type(e).__name__, e)}
    torch.cuda.empty_cache()
    try:
        out["mlp_fusion"] = time_prefill_mlp(bc, _cabi, dev, reps)
    except Exception as e:                                               # noqa: BLE001
        out["mlp_fusion"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:                                                                 # the same fusion at the default precision (fp32 spectra)
        out["mlp_fusion_fp32_spectra"] = time_prefill_mlp(bc, _cabi, dev, max(2, reps // 2), spectrum_dtype=torch.float32)
    except Exception as e:
        out["mlp_fusion_fp32_spectra"] = {"error": "%s: %s" % (type(e).__name__, e)}
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks of --steps steps; the median block is reported")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spectrum", default="float32", choices=["float32", "bfloat16"])
    ap.add_argument("--weight-sets", type=int, default=2)
    ap.add_argument("--ctas-per-sm", type=int, default=1)
    ap.add_argument("--no-detail", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
