"""Writes tests/golden/kat.json: the hand-derived known-answer vectors of SURVEY.md section 8(c).

The reference's own tests hold NO golden vectors (test/qr.jl, test/cholesky.jl and test/juliaBLAS.jl
assert properties only) and no Julia runtime exists in this image, so these vectors are derived by
hand from the published semantics of Julia's LinearAlgebra.reflector!/reflectorApply! and from the
reference source (src/qr.jl:64-111, src/cholesky.jl:37-55).  Exact rational arithmetic below; nothing
is computed by the oracle or by the GPU library."""
import json
from fractions import Fraction as F
import os

kat = {
    # A = [3 1; 4 2]: col 1 = (3,4): norm 5, nu = +5, xi = 8, R11 = -5, v2 = 4/8, tau1 = 8/5.
    # apply to col 2 = (1,2): s = tau (1 + 0.5*2) = 3.2 -> (1-3.2, 2-0.5*3.2) = (-2.2, 0.4);
    # col 2 tail = (0.4): length-1 vector still reflected: R22 = -0.4, tau2 = 2
    "qr_2x2": {"A": [[3.0, 1.0], [4.0, 2.0]], "factors": [[-5.0, -2.2], [0.5, -0.4]], "tau": [1.6, 2.0]},
    # zero column: tau = 0 and the vector is untouched
    "reflector_zero": {"x": [0.0, 0.0, 0.0], "x_out": [0.0, 0.0, 0.0], "tau": 0.0},
    # x = (-3, 4): nu = copysign(5, -3) = -5, xi = -8, x1 <- 5, x2 <- 4/(-8), tau = (-8)/(-5)
    "reflector_neg": {"x": [-3.0, 4.0], "x_out": [5.0, -0.5], "tau": 1.6},
    # complex x = (3i, 4): real(xi1) = 0 -> nu = +5, x1 <- -5, x2 <- 4/(5+3i), tau = (5+3i)/5
    "reflector_complex": {"x": [[0.0, 3.0], [4.0, 0.0]],
                          "x_out": [[-5.0, 0.0], [float(F(20, 34)), float(F(-12, 34))]],
                          "tau": [1.0, 0.6]},
    # A = [4 2; 2 5] -> in place [2 2; 1 2] (upper entry untouched), L = [2 0; 1 2]
    "chol_2x2": {"A": [[4.0, 2.0], [2.0, 5.0]], "inplace": [[2.0, 2.0], [1.0, 2.0]]},
    # T for a 2-column panel: T = [tau1, -tau1 tau2 (v1^H v2); 0, tau2]
    # V = [1 0; 0.5 1; 0.5 0.25], tau = (1.6, 1.25): v1^H v2 = 0.5*1 + 0.5*0.25 = 0.625
    "larft_2": {"F": [[9.0, 9.0], [0.5, 9.0], [0.5, 0.25]], "tau": [1.6, 1.25],
                "T": [[1.6, -1.6 * 1.25 * 0.625], [0.0, 1.25]]},
    # right apply, src/qr.jl:19-42: A = [1 2], x = (., 0.5), tau = 1.6: s = 1.6*(1+2*0.5)=3.2 -> (1-3.2, 2-3.2*0.5)
    "apply_right": {"A": [[1.0, 2.0]], "x": [123.0, 0.5], "tau": 1.6, "out": [[-2.2, 0.4]]},
    # rank-k generic, src/juliaBLAS.jl:89-112: C = I2, A = [1;2], alpha = -1 -> lower: [0 .; -2 -3], upper untouched
    "rank_update": {"C": [[1.0, 7.0], [0.0, 1.0]], "A": [[1.0], [2.0]], "alpha": -1.0,
                    "out": [[0.0, 7.0], [-2.0, -3.0]]},
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
    json.dump(kat, f, indent=1)
print("wrote kat.json")
