"""TEST INFRASTRUCTURE ONLY. CPU oracle for chad::TSDFMap::insert / Submap::finalize.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. Nothing under chad_tsdf_b200/ or include/ may import, link or execute it.
"""
