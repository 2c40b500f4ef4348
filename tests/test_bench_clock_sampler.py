"""bench.py's clock sampler (host logic; the contract wants SM clocks and throttle reasons sampled DURING the timed region,
which at 640x512 lasts ~20 ms): NVML thread with a fake pynvml, window filtering, and the no-NVML / no-nvidia-smi fallback."""
import sys
import time
import types


def _fake_nvml(reason_mask=0, init_ok=True):
    m = types.ModuleType("pynvml")
    m.NVML_CLOCK_SM = 1
    m.nvmlClocksEventReasonHwSlowdown, m.nvmlClocksEventReasonSwPowerCap = 0x8, 0x4
    m.nvmlClocksEventReasonSwThermalSlowdown, m.nvmlClocksEventReasonHwThermalSlowdown = 0x20, 0x40

    def init():
        if not init_ok:
            raise RuntimeError("NVML Shared Library Not Found")

    def by_uuid(u):
        if u not in ("GPU-abc", b"GPU-abc"):
            raise RuntimeError("not found")
        return "by-uuid"

    m.nvmlInit = init
    m.nvmlDeviceGetHandleByUUID = by_uuid
    m.nvmlDeviceGetHandleByIndex = lambda i: f"by-index-{i}"
    m.nvmlDeviceGetClockInfo = lambda h, c: 1965 if h == "by-uuid" else 1000
    m.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    m.nvmlDeviceGetCurrentClocksEventReasons = lambda h: reason_mask
    return m


def test_nvml_thread_reports_only_the_timed_window(monkeypatch):
    import bench

    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(reason_mask=0x4 | 0x40))
    s = bench.ClockSampler(3, "abc")  # torch reports the UUID without the "GPU-" prefix
    s.start()
    time.sleep(0.02)
    s.mark_begin()
    time.sleep(0.03)
    s.mark_end()
    time.sleep(0.01)
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0  # the handle found by UUID, not by index
    assert out["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"] and out["window"] == "timed region"
    assert 3 <= out["samples"] <= 20  # ~30 ms window at a ~2 ms period; samples outside it are not counted


def test_falls_back_without_nvml_and_without_nvidia_smi(monkeypatch):
    import bench

    monkeypatch.setitem(sys.modules, "pynvml", _fake_nvml(init_ok=False))
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0, "")
    s.start()
    s.mark_begin()
    s.mark_end()
    out = s.stop()
    assert out["sm_mhz"] is None and out["reasons"]


def test_roofline_row_uses_survey_bytes_and_drops_host_stall_samples():
    """bench._ka_row: algorithmic bytes = V * 4*B*H*W*(2C + D + G*D), mean launch time over the samples, and a
    sample that timed a host stall (far above the median) is dropped and counted instead of poisoning the mean."""
    import torch

    import bench

    B, H, W, C, V, D, G = 1, 64, 80, 64, 4, 32, 8
    ref, src, rt, depth = torch.zeros(B, H, W, C), torch.zeros(V, B, H, W, C), torch.zeros(V, B, 12), torch.zeros(B, D, H, W)
    times = [20e-6] * 9 + [60e-3]  # nine clean samples and one 60 ms host stall (run 17 of round 2 saw exactly this)
    row = bench._ka_row("warp_corr_score", (ref, src, rt, depth, G), {}, times, 6581.6)
    assert row["algorithmic_bytes"] == V * 4 * B * H * W * (2 * C + D + G * D) == 34078720
    assert row["dropped_host_stall_samples"] == 1 and row["samples"] == 9
    assert abs(row["us"] - 20.0) < 1e-6 and abs(row["frac"] - 34078720 / 20e-6 / 1e9 / 6581.6) < 1e-9
    clean = bench._ka_row("warp_corr_score", (ref, src, rt, depth, G), {}, [20e-6, 22e-6, 21e-6], 6581.6)
    assert clean["dropped_host_stall_samples"] == 0 and clean["samples"] == 3
