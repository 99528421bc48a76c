"""Test infrastructure (the CPU oracle and the compiled reference).  Never imported by qpalm_b200/."""
