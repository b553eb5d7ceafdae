"""Writes tests/golden/reference_goldens.json.

The reference (TemGym/TemGymCore) cannot be imported in this environment (jax is not
installed, src/temgym_core/__init__.py:4), so its golden vectors are TRANSCRIBED here from
the reference's own README, notebooks' stored cell outputs and tests, each with the
file:line (or notebook cell) it was read from.  Nothing here is computed by our code.
Run:  python tests/golden/make_reference_goldens.py
"""
import json
import os

G = {}

# README.md:35-52 -- quick start
G["readme_ray"] = {
    "cite": "README.md:35-52",
    "ray_in": dict(x=0.1, y=0.2, dx=0.3, dy=0.4, z=0.0, pathlength=0.0),
    "model": [["Lens", dict(z=0.5, focal_length=1.0)],
              ["Detector", dict(z=1.0, pixel_size=[0.01, 0.01], shape=[128, 128])]],
    # printed: Ray(x=0.275, y=0.4, dx=0.05, dy=0.0, z=1.0, pathlength=0.89) (2 s.f. print of 0.88875)
    "ray_out_printed": dict(x=0.275, y=0.4, dx=0.05, dy=0.0, z=1.0, pathlength=0.89),
    "print_decimals": 3,
}
# README.md:227-235
G["readme_abcd"] = {
    "cite": "README.md:227-235",
    "abcd": [[0.5, 0.0, 0.75, 0.0, 0.0], [0.0, 0.5, 0.0, 0.75, 0.0], [-1.0, 0.0, 0.5, 0.0, 0.0],
             [0.0, -1.0, 0.0, 0.5, 0.0], [0.0, 0.0, 0.0, 0.0, 1.0]],
}
# README.md:240-268 -- solve_model per-step matrices
_P = [[1.0, 0, 0.5, 0, 0], [0, 1.0, 0, 0.5, 0], [0, 0, 1.0, 0, 0], [0, 0, 0, 1.0, 0], [0, 0, 0, 0, 1.0]]
_L = [[1.0, 0, 0, 0, 0], [0, 1.0, 0, 0, 0], [-1.0, 0, 1.0, 0, 0], [0, -1.0, 0, 1.0, 0], [0, 0, 0, 0, 1.0]]
_I = [[1.0 if i == j else 0.0 for j in range(5)] for i in range(5)]
G["readme_solve_model"] = {"cite": "README.md:240-268", "per_step": [_P, _L, _P, _I]}
# README.md:147-160 -- gradients wrt the input ray
G["readme_input_grads"] = {
    "cite": "README.md:141-160",
    "ray_in": dict(x=0.01, y=0.2, dx=0.3, dy=0.4, z=0.0, pathlength=0.6),
    "d_dx_out_d_x_in": -1.0,
    "d_dy_out_d_x_in": 0.0,
}
# README.md:120-137 -- gradients wrt lens parameters (f, z): printed (0.125, 0.09999999)
G["readme_param_grads"] = {
    "cite": "README.md:120-137",
    "ray_in": dict(x=0.1, y=0.2, dx=0.3, dy=0.4, z=0.0, pathlength=0.0),
    "d_x_out_d_f": 0.125, "d_x_out_d_z": 0.1, "print_rtol": 1e-6,
}
# examples/aperture_diffraction.ipynb cell 13 (stored output)
G["aperture_diffraction_abcd"] = {
    "cite": "examples/aperture_diffraction.ipynb cells 11,13",
    "F1": 1.0e-2,
    "model": [["ParallelBeam", dict(z=0.0, radius=1e-7)],
              ["Lens", dict(z=1.0e-2, focal_length=1.0e-2)],
              ["Detector", dict(z=2.0e-2, pixel_size=[2e-8, 2e-8], shape=[512, 512])]],
    "ray_in": dict(x=0.0, y=0.0, dx=0.0, dy=0.0, z=0.0, pathlength=0.0),
    "abcd": [[0.0, 0.0, 1e-2, 0.0, 0.0], [0.0, 0.0, 0.0, 1e-2, 0.0], [-1e2, 0.0, 0.0, 0.0, 0.0],
             [0.0, -1e2, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0, 1.0]],
    "rtol": 1e-12, "atol": 1e-15,
}
# examples/two_beam_interference.ipynb cells 7,8 (stored output, 7 significant figures)
G["two_beam_abcd"] = {
    "cite": "examples/two_beam_interference.ipynb cells 7,8",
    "params": dict(scale=1e6, M1=-100, F1=5e-3 * 1e6, defocus=1e-10 * 1e6, def_x=-0.4e-4,
                   pixel_size=0.1e-6 * 1e6),
    "ray_in": dict(x=1e-12, y=0.0, dx=0.0, dy=0.0, pathlength=0.0),
    "abcd": [[-1.999900e+02, 0.0, -4.999520e+03, 0.0, 1.010000e+01],
             [0.0, -1.999900e+02, 0.0, -4.999520e+03, 0.0],
             [-2.000000e-04, 0.0, -1.000002e-02, 0.0, 4.000000e-05],
             [0.0, -2.000000e-04, 0.0, -1.000002e-02, 0.0],
             [0.0, 0.0, 0.0, 0.0, 1.0]],
    "rtol": 1e-6,
}
# examples/biprism.ipynb cells 6,7 (stored output, 9 significant figures)
G["biprism_abcd"] = {
    "cite": "examples/biprism.ipynb cells 6,7",
    "params": dict(M1=-200, F1=0.0025, M2=-1500, F2=0.02, defocus=1e-9, def_x=-2e-5,
                   aperture_radius=50e-9),
    "ray_in": dict(x=1e-15, y=0.0, dx=0.0, dy=0.0, pathlength=0.0),
    "abcd": [[3.00000000e+05, 0.0, -3.00000000e-04, 0.0, -7.53750000e-03],
             [0.0, 3.00000000e+05, 0.0, -3.00000000e-04, 0.0],
             [1.00002667e+04, 0.0, -6.66693333e-06, 0.0, -2.51263333e-04],
             [0.0, 1.00002667e+04, 0.0, -6.66693333e-06, 0.0],
             [0.0, 0.0, 0.0, 0.0, 1.0]],
    "rtol": 2e-8,
}
# tests/test_component.py:426-465 with tests/transfer_matrices.py:110-231 (atol 1e-12)
G["biprism_lens_prop"] = {
    "cite": "tests/test_component.py:426-465; tests/transfer_matrices.py:110-231",
    "params": dict(M1=-10, F1=0.0002, defocus=1e-4, deflection=1e-4),
    "atol": 1e-12,
}
# tests/test_component.py:73-174 -- grid tables (11x11, pixel 0.1)
G["grid_tables"] = {
    "cite": "tests/test_component.py:73-174",
    "shape": [11, 11], "pixel_size": [0.1, 0.1],
    "m2p": [  # (xy, rotation, expected (py, px))
        [[0.0, 0.0], 0.0, [5, 5]], [[-0.5, 0.5], 0.0, [0, 0]], [[0.5, -0.5], 0.0, [10, 10]],
        [[0.0, 0.5], 0.0, [0, 5]], [[-0.5, 0.0], 0.0, [5, 0]],
        [[0.0, 0.0], 90.0, [5, 5]], [[-0.5, 0.5], 90.0, [10, 0]], [[0.5, -0.5], 90.0, [0, 10]],
        [[0.0, 0.5], 90.0, [5, 0]], [[-0.5, 0.0], 90.0, [10, 5]],
    ],
    "p2m": [  # (pixel (py,px), rotation, expected (x, y))
        [[5, 5], 0.0, [0.0, 0.0]], [[0, 0], 0.0, [-0.5, 0.5]], [[10, 10], 0.0, [0.5, -0.5]],
        [[0, 5], 0.0, [0.0, 0.5]], [[5, 0], 0.0, [-0.5, 0.0]],
        [[5, 5], 90.0, [0.0, 0.0]], [[10, 0], 90.0, [-0.5, 0.5]], [[0, 10], 90.0, [0.5, -0.5]],
        [[5, 0], 90.0, [0.0, 0.5]], [[10, 5], 90.0, [-0.5, 0.0]],
    ],
    "atol": 1e-6,
}
# tests/test_gaussians.py:229-273 -- free-space field KAT, rtol 1e-9 / atol 1e-12
G["free_space_field_kat"] = {
    "cite": "tests/test_gaussians.py:229-273",
    "w0": 1e-3, "wl": 500e-9, "L": 0.25,
    "xs": [-1e-4, 0.0, 2e-4], "ys": [0.0, 1e-4, -2e-4],
    "rtol": 1e-9, "atol": 1e-12,
}
# tests/test_gaussians.py:218-226
G["qinv_identity"] = {"cite": "tests/test_gaussians.py:218-226", "qx_im": 2.0, "qy_im": 3.0,
                      "rtol": 1e-12}
# tests/test_component.py:398-423 -- biprism Jacobian
G["biprism_jac"] = {"cite": "tests/test_component.py:398-423", "deflection": 1e-3, "z_det": 0.234,
                    "atol": 1e-6}

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
with open(out, "w") as fh:
    json.dump(G, fh, indent=1)
print("wrote", out)
