"""Synthetic genomes for BASELINE.json configs C1..C5: re-export of tools/synth (own host library libmbsynth.so;
the generator is bench / test infrastructure and is not part of libmauve_b200.so)."""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "synth", "__init__.py")
_spec = importlib.util.spec_from_file_location("mb_synth_tools", _path)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
synth_genomes = _mod.synth_genomes
