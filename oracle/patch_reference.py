#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Makes a COPY of two reference source files with the two call-site hooks of
INTEGRATION.md inserted (the reference tree itself is read-only and is never modified or copied
into the repo):

    python oracle/patch_reference.py <reference>/src <outdir>

writes <outdir>/force/force.cu and <outdir>/integrate/integrate.cu.  The edits are anchored on short
unique strings of the reference so that a reference update that moves them fails loudly here."""
import sys
from pathlib import Path


def insert_after(text, anchor, addition, what):
    if text.count(anchor) != 1:
        raise SystemExit(f"patch_reference: anchor for {what} found {text.count(anchor)} times: {anchor!r}")
    return text.replace(anchor, anchor + addition)


def insert_before(text, anchor, addition, what):
    if text.count(anchor) != 1:
        raise SystemExit(f"patch_reference: anchor for {what} found {text.count(anchor)} times: {anchor!r}")
    return text.replace(anchor, addition + anchor)


def main():
    src, out = Path(sys.argv[1]), Path(sys.argv[2])
    # ---- Force::parse_potential: ask libb200md first
    f = (src / "force" / "force.cu").read_text()
    f = insert_after(f, '#include "force.cuh"\n', '#include "b200md_gpumd.cuh"\n', "force.cu include")
    f = insert_after(
        f, "  bool is_nep = false;\n",
        "  if (b200md_make_potential(potential_name, param[1], number_of_atoms, potential)) {\n"
        "    is_nep = strncmp(potential_name, \"nep\", 3) == 0;\n"
        "    if (is_nep)\n"
        "      check_types(param[1]);\n"
        "  } else\n", "parse_potential hook")
    (out / "force").mkdir(parents=True, exist_ok=True)
    (out / "force" / "force.cu").write_text(f)
    # ---- Integrate::initialize: ask libb200md first
    g = (src / "integrate" / "integrate.cu").read_text()
    g = insert_after(g, '#include "integrate.cuh"\n', '#include "b200md_gpumd.cuh"\n', "integrate.cu include")
    g = insert_before(
        g, "  switch (type) {\n    case 0: // NVE\n",
        "  if (!b200md_make_ensemble(\n"
        "        type, move_group, move_velocity, number_of_atoms, temperature, temperature_coupling,\n"
        "        time_step, target_pressure, num_target_pressure_components, pressure_coupling,\n"
        "        deform_x, deform_y, deform_z, deform_rate, ensemble))\n", "Integrate::initialize hook")
    (out / "integrate").mkdir(parents=True, exist_ok=True)
    (out / "integrate" / "integrate.cu").write_text(g)
    print("patched copies written to", out)


if __name__ == "__main__":
    main()
