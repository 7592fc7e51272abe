"""Import shim: the package directory is named `flight.jl_b200` (not a valid Python identifier), so this
module loads it under the name `flightbatch`. `import flightbatch as fb` is the public entry point."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flight.jl_b200")
_spec = importlib.util.spec_from_file_location("flightbatch", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["flightbatch"] = _mod
_spec.loader.exec_module(_mod)
