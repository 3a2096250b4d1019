"""TEST INFRASTRUCTURE: cosma_b200_plan_host_panel through ctypes -> (eligible, b_pieces, c_pieces), pieces = [(src_off, len, dst_off)]."""
import ctypes


def host_panel(lib, plan_handle, c, j, cap=3000):
    bp, cp = (ctypes.c_int64 * cap)(), (ctypes.c_int64 * cap)()
    nb, nc, ok = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    st = lib.cosma_b200_plan_host_panel(plan_handle, c, j, bp, cap, ctypes.byref(nb), cp, cap, ctypes.byref(nc), ctypes.byref(ok))
    assert st == 0
    return bool(ok.value), [(bp[3 * i], bp[3 * i + 1], bp[3 * i + 2]) for i in range(nb.value)], [(cp[3 * i], cp[3 * i + 1], cp[3 * i + 2]) for i in range(nc.value)]
