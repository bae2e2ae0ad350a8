"""CPU oracle for the ClusterCRF hot path — TEST INFRASTRUCTURE, never imported by gecco_b200/.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` leg
may use this package (as the checker and the CPU timing baseline).
"""
