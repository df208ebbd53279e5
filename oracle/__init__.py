"""CPU restatement of the reference's explicit dGSEM path -- TEST INFRASTRUCTURE ONLY.

Nothing under nebulasem_b200/ imports, links or executes anything in this directory; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs do, and there only as the checker.

  mesh.py        topology + element geometry (addBoundaryCells, fixHexCells, calcGeometry incl. the cubed-sphere corrections, ExtrudeMesh)
  dg.py          basis, node placement (incl. the radial rescale on the sphere), face maps, Jinv, mortar projections
  euler.py       the euler app's set-up and time-loop body, the operators (cds / uds / rusanov / gradf / divf / BCs)
  convection.py  the convection app (winds NONE / LEVEQUE / LAURITZEN_0/1, AB1-AB5, RUSANOV / CDS / UDS / BLENDED)
  amr.py         MeshField::refineField (copy / refine / coarsen with the mass fix)
  case.py        a case directory -> oracle (controls, grid, field files with their initialisers)
  cases.py       synthetic cases;  refio.py  the reference's file formats;  libm.py  the platform libm through ctypes (csrc/libm_exact.c)
  _ref/          the UNMODIFIED reference compiled where it lies (build_ref.sh, serial MPI shim in mpi_shim/), with two own drivers linked
                 against its objects (tools/geomdump.cpp, tools/refinedump.cpp)

Parity PINNED: every module is checked against fixtures written by the unmodified reference binaries (tests/golden/, each with the script
that made it) -- geometry arrays bit for bit (flat, non-conforming, cubed sphere, the reference's own regridded sphere), step dumps of all
the reference's DG euler and convection examples (bit-identical with exact_order=True), 100-step runs, AMR field transfers of four regrids --
and, where oracle/_ref is present, against the reference run live (tests/test_oracle_vs_reference.py).
"""
