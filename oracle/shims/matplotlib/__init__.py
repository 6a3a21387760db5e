"""Empty stand-in for matplotlib (imported at F16model/env/task.py:2 and
F16model/utils/plots.py:3; never called on the hot path). TEST INFRASTRUCTURE ONLY."""
