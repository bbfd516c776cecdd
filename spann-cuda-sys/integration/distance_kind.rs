// Goes into src/distances/distance.rs (reference lines 7-43).  UNCOMPILED.
// The per-pair trait stays (API compatibility); `kind()` tells batched callers which metric an
// `Arc<dyn DistanceMetric<F>>` is, so that they can hand the whole batch to the GPU library.
use spann_cuda_sys::MetricKind;

pub trait DistanceMetric<F: Float + Debug + Default>: Debug + Send + Sync {
    fn compute(&self, a: &ArrayView1<F>, b: &ArrayView1<F>) -> F;
    /// `None`: a user-defined metric, only the CPU path applies.
    fn kind(&self) -> Option<MetricKind> {
        None
    }
}

// in the three impl blocks (lines 16-21, 27-32, 38-43) add respectively:
//     fn kind(&self) -> Option<MetricKind> { Some(MetricKind::Euclidean) }
//     fn kind(&self) -> Option<MetricKind> { Some(MetricKind::Manhattan) }
//     fn kind(&self) -> Option<MetricKind> { Some(MetricKind::Chebyshev) }
