"""CPU oracle of the VolumetricReSTIR hot path — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline)."""
