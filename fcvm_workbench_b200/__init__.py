"""fcvm_workbench_b200 -- B200-native Newton-Raphson load-stepping path of fcVM."""
__version__ = "0.1.0"
