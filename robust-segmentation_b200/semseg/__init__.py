"""Host-side mirror of the reference's ``semseg`` call surface for the attack-side hot path.

Only the modules on that path exist here: ``attacker``, ``losses``, ``metrics``, ``val``.
Models, datasets, optimisers and schedulers stay with the reference.
"""
