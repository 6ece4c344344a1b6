"""CPU oracle for the libstacker ECC align-and-stack path.  TEST INFRASTRUCTURE ONLY:
imported by tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke();
never by the product package libstacker.rs_b200/."""
