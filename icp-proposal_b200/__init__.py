"""icp-proposal_b200: B200-native hot path of unibas-gravis/icp-proposal (libicpcuda.so + host mirror)."""
__version__ = "0.1.0"
