"""Writes tests/golden/costa_copy_and_transform.json: the known-answer cases of the reference's COSTA unit test
(libs/COSTA/tests/unit/test_utils.cpp:7-270) for copy_and_transform -- argument lists, input arrays and expected
outputs. The fourth case (1000 x 500 col->row, srand(100)) is regenerated from its rule `in[i] = i + rand()` with the C
library's rand() and its expected output is defined by the test's own predicate out[i*ld_out + j] == in[j*ld_in + i].

    python tests/golden/make_costa_golden.py
"""
import ctypes
import json
import os

IN = [9, 1, 1, -1, 7, 3, 4, -1, 5, 5, 1, -1, 9, 2, 3, -1, 7, 6, 5, -1, 2, 2, 4, -1, 3, 7, 4, -1, 3, 8, 1, -1]


def main():
    cases = []
    # copy2D row_major (test_utils.cpp:7-73) and col_major (:75-141): strided copy 4 -> 5
    copy_result = []
    for r in range(8):
        copy_result += IN[4 * r:4 * r + 4] + [-1]
    for name, n_rows, n_cols, so, do in (("copy2D_row_major", 8, 3, "R", "R"), ("copy2D_col_major", 3, 8, "C", "C")):
        cases.append({"name": name, "n_rows": n_rows, "n_cols": n_cols, "src_ld": 4, "dst_ld": 5, "src_ordering": so, "dst_ordering": do,
                      "transpose": 0, "conjugate": 0, "alpha": 1, "beta": 0, "in": IN, "expected": copy_result,
                      "compare": "logical"})
    # transpose row_to_col_major (:143-206)
    cases.append({"name": "transpose_row_to_col_major", "n_rows": 8, "n_cols": 3, "src_ld": 4, "dst_ld": 10, "src_ordering": "R",
                  "dst_ordering": "C", "transpose": 0, "conjugate": 0, "alpha": 1, "beta": 0, "in": IN,
                  "expected": [9, 7, 5, 9, 7, 2, 3, 3, -1, -1, 1, 3, 5, 2, 6, 2, 7, 8, -1, -1, 1, 4, 1, 3, 5, 4, 4, 1, -1, -1],
                  "compare": "logical"})
    # transpose col_to_row_major (:208-270): 1000 x 500, strides 1100 -> 501, in[i] = i + rand() after srand(100)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(100)
    n_rows, n_cols, ld_in, ld_out = 1000, 500, 1100, 501
    big = [(i + libc.rand()) & 0xFFFFFFFF for i in range(n_cols * ld_in)]
    big = [v - (1 << 32) if v >= (1 << 31) else v for v in big]  # int wrap-around like the C test
    cases.append({"name": "transpose_col_to_row_major", "n_rows": n_rows, "n_cols": n_cols, "src_ld": ld_in, "dst_ld": ld_out,
                  "src_ordering": "C", "dst_ordering": "R", "transpose": 0, "conjugate": 0, "alpha": 1, "beta": 0,
                  "in_rule": "in[i] = (int)(i + rand()) after srand(100), i < n_cols*src_ld", "in_checksum": sum(big) & 0xFFFFFFFFFFFF,
                  "in_head": big[:16], "compare": "predicate out[i*dst_ld + j] == in[j*src_ld + i]"})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "costa_copy_and_transform.json")
    with open(out, "w") as f:
        json.dump({"source": "reference libs/COSTA/tests/unit/test_utils.cpp", "cases": cases}, f)
    print(out, len(cases))


if __name__ == "__main__":
    main()
