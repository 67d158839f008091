"""Certificate types of the operator contract: every operator returns (result, CertBundle, ExpectedEffect)
(docs/OPERATOR_CONTRACTS.md:3).  Field-for-field restatement of the parts of pkg/common/certificates.py the hot path
fills (:21-108 component certs, :348-503 bundle + effect), so ``to_dict()`` output is interchangeable with the
reference's (checked against tests/golden/ref_certs.json).  The OT / map-update / compute sub-certs of the
reference are not produced by this path and are carried only as their default dictionaries.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any, Dict, List, Optional


@dataclass
class ConditioningCert:
    eig_min: float = 1.0
    eig_max: float = 1.0
    cond: float = 1.0
    near_null_count: int = 0

    def to_dict(self) -> Dict[str, Any]:
        return {"eig_min": self.eig_min, "eig_max": self.eig_max, "cond": self.cond,
                "near_null_count": self.near_null_count}


@dataclass
class SupportCert:
    ess_total: float = 0.0
    support_frac: float = 1.0

    def to_dict(self) -> Dict[str, Any]:
        return {"ess_total": self.ess_total, "support_frac": self.support_frac}


@dataclass
class MismatchCert:
    nll_per_ess: float = 0.0
    directional_score: float = 1.0

    def to_dict(self) -> Dict[str, Any]:
        return {"nll_per_ess": self.nll_per_ess, "directional_score": self.directional_score}


@dataclass
class ExcitationCert:
    dt_effect: float = 0.0
    extrinsic_effect: float = 0.0

    def to_dict(self) -> Dict[str, Any]:
        return {"dt_effect": self.dt_effect, "extrinsic_effect": self.extrinsic_effect}


@dataclass
class InfluenceCert:
    lift_strength: float = 0.0
    psd_projection_delta: float = 0.0
    nu_projection_delta: float = 0.0
    mass_epsilon_ratio: float = 0.0
    anchor_drift_rho: float = 0.0
    dt_scale: float = 1.0
    extrinsic_scale: float = 1.0
    trust_alpha: float = 1.0
    power_beta: float = 1.0

    @classmethod
    def identity(cls) -> "InfluenceCert":
        return cls()

    def with_overrides(self, **kwargs: Any) -> "InfluenceCert":
        return replace(self, **kwargs)

    def to_dict(self) -> Dict[str, Any]:
        return {k: getattr(self, k) for k in ("lift_strength", "psd_projection_delta", "nu_projection_delta",
                                              "mass_epsilon_ratio", "anchor_drift_rho", "dt_scale", "extrinsic_scale",
                                              "trust_alpha", "power_beta")}


@dataclass
class DeviceRuntimeCert:
    """Host/device transfer accounting (pkg/common/certificates.py:299-315)."""
    host_sync_count_est: int = 0
    device_to_host_bytes_est: int = 0
    host_to_device_bytes_est: int = 0
    jit_recompile_count: int = 0

    def to_dict(self) -> Dict[str, Any]:
        return {k: int(getattr(self, k)) for k in ("host_sync_count_est", "device_to_host_bytes_est",
                                                   "host_to_device_bytes_est", "jit_recompile_count")}


@dataclass
class ComputeCert:
    """Shape / allocation accounting (pkg/common/certificates.py:317-345)."""
    alloc_bytes_est: int = 0
    largest_tensor_shape: tuple = (0, 0)
    segment_sum_k: int = 0
    psd_projection_count: int = 0
    chol_solve_count: int = 0
    device_runtime: DeviceRuntimeCert = field(default_factory=DeviceRuntimeCert)

    def to_dict(self) -> Dict[str, Any]:
        return {"alloc_bytes_est": int(self.alloc_bytes_est), "largest_tensor_shape": self.largest_tensor_shape,
                "segment_sum_k": int(self.segment_sum_k), "psd_projection_count": int(self.psd_projection_count),
                "chol_solve_count": int(self.chol_solve_count),
                "scan_io": {"scan_seq": 0, "scan_stamp_sec": 0.0, "scan_window_start_sec": 0.0,
                            "scan_window_end_sec": 0.0, "streams": {}},
                "device_runtime": self.device_runtime.to_dict()}


_OVERCONFIDENCE_DEFAULT = {"excitation_total": 0.0, "ess_to_excitation": 0.0, "cond_to_support": 0.0, "dt_asymmetry": 0.0,
                           "z_to_xy_ratio": 0.0, "ess_growth_rate": 0.0, "excitation_growth_rate": 0.0,
                           "nullspace_energy_ratio": 0.0}


@dataclass
class CertBundle:
    chart_id: str
    anchor_id: str
    exact: bool
    approximation_triggers: List[str] = field(default_factory=list)
    frobenius_applied: bool = False
    conditioning: ConditioningCert = field(default_factory=ConditioningCert)
    support: SupportCert = field(default_factory=SupportCert)
    mismatch: MismatchCert = field(default_factory=MismatchCert)
    excitation: ExcitationCert = field(default_factory=ExcitationCert)
    influence: InfluenceCert = field(default_factory=InfluenceCert)
    compute: ComputeCert = field(default_factory=ComputeCert)

    @classmethod
    def create_exact(cls, chart_id: str, anchor_id: str, conditioning: Optional[ConditioningCert] = None,
                     support: Optional[SupportCert] = None, mismatch: Optional[MismatchCert] = None,
                     excitation: Optional[ExcitationCert] = None, influence: Optional[InfluenceCert] = None,
                     compute: Optional[ComputeCert] = None) -> "CertBundle":
        return cls(chart_id=chart_id, anchor_id=anchor_id, exact=True, approximation_triggers=[], frobenius_applied=False,
                   conditioning=conditioning or ConditioningCert(), support=support or SupportCert(),
                   mismatch=mismatch or MismatchCert(), excitation=excitation or ExcitationCert(),
                   influence=influence or InfluenceCert(), compute=compute or ComputeCert())

    @classmethod
    def create_approx(cls, chart_id: str, anchor_id: str, triggers: List[str], frobenius_applied: bool = False,
                      conditioning: Optional[ConditioningCert] = None, support: Optional[SupportCert] = None,
                      mismatch: Optional[MismatchCert] = None, excitation: Optional[ExcitationCert] = None,
                      influence: Optional[InfluenceCert] = None, compute: Optional[ComputeCert] = None) -> "CertBundle":
        return cls(chart_id=chart_id, anchor_id=anchor_id, exact=False, approximation_triggers=list(triggers),
                   frobenius_applied=frobenius_applied, conditioning=conditioning or ConditioningCert(),
                   support=support or SupportCert(), mismatch=mismatch or MismatchCert(),
                   excitation=excitation or ExcitationCert(), influence=influence or InfluenceCert(),
                   compute=compute or ComputeCert())

    def total_trigger_magnitude(self) -> float:
        """What the downstream Frobenius step consumes (pkg/common/certificates.py:439-455, pipeline.py:1211)."""
        i = self.influence
        return (i.lift_strength + i.psd_projection_delta + i.nu_projection_delta + i.mass_epsilon_ratio
                + i.anchor_drift_rho + abs(1.0 - i.dt_scale) + abs(1.0 - i.extrinsic_scale) + abs(1.0 - i.trust_alpha)
                + abs(1.0 - i.power_beta))

    def to_dict(self) -> Dict[str, Any]:
        return {"chart_id": self.chart_id, "anchor_id": self.anchor_id, "exact": self.exact,
                "approximation_triggers": self.approximation_triggers, "frobenius_applied": self.frobenius_applied,
                "conditioning": self.conditioning.to_dict(), "support": self.support.to_dict(),
                "mismatch": self.mismatch.to_dict(), "excitation": self.excitation.to_dict(),
                "influence": self.influence.to_dict(), "overconfidence": dict(_OVERCONFIDENCE_DEFAULT),
                "compute": self.compute.to_dict(), "total_trigger_magnitude": self.total_trigger_magnitude()}


@dataclass
class ExpectedEffect:
    objective_name: str
    predicted: float
    realized: Optional[float] = None

    def to_dict(self) -> Dict[str, Any]:
        return {"objective_name": self.objective_name, "predicted": self.predicted, "realized": self.realized}
