"""Copies the numbers of the reference's Taylor-Green vortex test outputs into
tests/golden/reference_golden.json (run in the build container, where /root/reference exists):

    python tests/golden/extract_taylor_green.py [/root/reference]

applications_tests/gls_navier_stokes_2d/taylor-green-vortex_gls_{sdirk3,sdirk2,bdf1}.mpirun=2.output:
periodic [0, 2 pi]^2, nu = 1, L2-projected initial condition; per transient iteration the CFL number
(of the solution at the start of the step), enstrophy, kinetic energy and the velocity L2 error are
printed.  Numbers are kept as the strings the reference printed (digit-for-digit comparison)."""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))
path = os.path.join(here, "reference_golden.json")
with open(path) as f:
    gold = json.load(f)

for scheme in ("sdirk3", "sdirk2", "bdf1"):
    rel = "applications_tests/gls_navier_stokes_2d/taylor-green-vortex_gls_%s.mpirun=2.output" % scheme
    text = open(os.path.join(ref, rel)).read()
    head, _, rest = text.partition("*****")
    entry = {
        "source": rel,
        "prm": rel.replace(".mpirun=2.output", ".prm"),
        "cells": int(re.search(r"Number of active cells:\s+(\d+)", head).group(1)),
        "dofs": int(re.search(r"Number of degrees of freedom:\s+(\d+)", head).group(1)),
        "enstrophy_0": re.search(r"Enstrophy\s+:\s+(\S+)", head).group(1),
        "kinetic_energy_0": re.search(r"Kinetic energy\s+:\s+(\S+)", head).group(1),
        "cfl": re.findall(r"CFL : (\S+)", rest),
        "enstrophy": re.findall(r"Enstrophy\s+:\s+(\S+)", rest),
        "kinetic_energy": re.findall(r"Kinetic energy\s+:\s+(\S+)", rest),
        "l2_error_velocity": re.findall(r"L2 error velocity : (\S+)", rest),
        "error_table": re.findall(r"^(\d\.\d{4}) (\d\.\d{4}e[-+]\d\d)", rest, flags=re.M),
    }
    gold["taylor_green_vortex_" + scheme] = entry
    print(scheme, entry["cells"], entry["dofs"], len(entry["cfl"]), len(entry["error_table"]))

with open(path, "w") as f:
    json.dump(gold, f, indent=1)
