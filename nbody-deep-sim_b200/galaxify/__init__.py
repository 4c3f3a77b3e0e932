"""galaxify on B200: drop-in for the reference package of the same name (src/galaxify/ in bikuta6/nbody-deep-sim).

`from galaxify import galaxies, simulation` works as in src/s01-dataset-generation.py:8 once this package's parent
directory (nbody-deep-sim_b200/) is on sys.path, which is how the reference relies on src/ being sys.path[0].
"""
