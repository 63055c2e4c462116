"""Plain configuration structs: mirror of plonky2/src/plonk/{config,circuit_data}.rs (API surface only)."""
from dataclasses import dataclass, field

from .fri import FriConfig, FriReductionStrategy
from .hashing import PoseidonHash


class PoseidonGoldilocksConfig:
    """plonk/config.rs:101-108: F = GoldilocksField, FE = quadratic extension (D = 2),
    Hasher = InnerHasher = PoseidonHash."""

    D = 2
    Hasher = PoseidonHash
    InnerHasher = PoseidonHash


@dataclass
class CircuitConfig:
    """plonk/circuit_data.rs:41-59"""

    num_wires: int = 135
    num_routed_wires: int = 80
    num_constants: int = 2
    use_base_arithmetic_gate: bool = True
    security_bits: int = 100
    num_challenges: int = 2
    zero_knowledge: bool = False
    max_quotient_degree_factor: int = 8
    fri_config: FriConfig = field(
        default_factory=lambda: FriConfig(
            rate_bits=3,
            cap_height=4,
            proof_of_work_bits=16,
            reduction_strategy=FriReductionStrategy.ConstantArityBits(4, 5),
            num_query_rounds=28,
        )
    )

    @classmethod
    def standard_recursion_config(cls):
        """plonk/circuit_data.rs:72-91"""
        return cls()

    def num_advice_wires(self):
        return self.num_wires - self.num_routed_wires
