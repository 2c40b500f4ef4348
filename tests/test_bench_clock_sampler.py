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
