"""Placeholder so `import tensorflow_addons` in the reference's optimiser code does
not fail if reached; nothing on the hot path uses it.  TEST INFRASTRUCTURE ONLY."""
