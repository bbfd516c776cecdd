// Goes into src/clustering/hierarchical.rs.  UNCOMPILED.  F = f32 (all BASELINE configs); a generic
// F keeps the CPU path.  `gpu: Option<GpuState>` is a new private field of HierarchicalClustering,
// filled by `new` when every row is contiguous and the metric has a `kind()`.
use spann_cuda_sys::{Assignment, Context, Dataset, MetricKind};

struct GpuState<'a> {
    ctx: &'a Context,
    ds: Dataset<'a>,
    kind: MetricKind,
    last: Option<Assignment>,        // assignment of the last assign_points (update_centroids reuses it)
}

impl<'a> HierarchicalClustering<'a, f32> {
    // reference :57-63 — additionally upload the borrowed rows once
    fn attach_gpu(&mut self, ctx: &'a Context) -> Result<(), Box<dyn Error>> {
        if let Some(kind) = self.params.distance_metric.kind() {
            let ds = Dataset::upload(ctx, &self.embeddings)?;
            self.gpu = Some(GpuState { ctx, ds, kind, last: None });
        }
        Ok(())
    }

    // reference :295-364 — same signature, same result (cluster-major lists, input order)
    fn assign_points_to_clusters(&self, point_indices: &[usize], centroids: &[(usize, usize)]) -> Vec<Vec<usize>> {
        let g = self.gpu.as_ref().expect("GPU state");
        let rows: Vec<usize> = centroids.iter().map(|&(row, _depth)| row).collect();
        let whole = point_indices.len() == self.embeddings.nrows() && point_indices.iter().enumerate().all(|(i, &p)| i == p);
        let a = g.ds.assign(g.kind, if whole { None } else { Some(point_indices) }, &rows, BOUNDARY_THRESHOLD as f32)
                    .expect("spf_assign");
        a.cluster_lists().expect("spf_assign_fetch")
    }

    // reference :138-181 — medoid of every cluster; an empty cluster keeps its centroid
    fn update_centroids(&mut self) {
        let g = self.gpu.as_ref().expect("GPU state");
        let old: Vec<usize> = self.clusters.iter().map(|c| c.centroid_idx.unwrap_or(0)).collect();
        let a = g.last.as_ref().expect("assign_points ran before update_centroids");
        let new = g.ds.update_medoids(g.kind, a, &old).expect("spf_update_medoids_from");
        for (c, row) in self.clusters.iter_mut().zip(new) {
            c.centroid_idx = Some(row);
        }
    }

    // reference :107-126 — c1 by rng, c2 = farthest member (strict >, identity (0, 0.0))
    fn bisect_seeds(&self, points: &[usize], rng: &mut impl Rng) -> (usize, usize) {
        let g = self.gpu.as_ref().expect("GPU state");
        let c1 = *points.choose(rng).expect("non-empty cluster");
        let c2 = g.ds.farthest(g.kind, c1, points).expect("spf_farthest");
        (c1, c2)
    }

    // reference :249-293 — the host keeps the RNG; up to 256 rounds per library call (spf_kmpp_rounds:
    // the draws are made ahead, the rounds run back to back on the device, one synchronisation)
    fn initialize_clusters_kmeans_plus_plus(&mut self, rng: &mut impl Rng) {
        use spann_cuda_sys::ffi;
        let g = self.gpu.as_ref().expect("GPU state");
        let n = self.embeddings.nrows();
        let first = (0..n).choose(rng).expect("non-empty data");
        let mut rows = vec![first];
        let mut s = std::ptr::null_mut();
        unsafe { ffi::spf_kmpp_begin(g.ds.raw(), g.kind as i32, first as u64, &mut s) };
        let mut draws: Vec<f64> = Vec::new();
        let mut picked = vec![0u64; 256];
        while rows.len() < self.params.initial_k {
            let want = (self.params.initial_k - rows.len()).min(256);
            while draws.len() < want { draws.push(rng.random::<f64>()); }
            let mut done = 0u32;
            let rc = unsafe { ffi::spf_kmpp_rounds(s, draws.as_ptr(), want as u32, picked.as_mut_ptr(), &mut done) };
            rows.extend(picked[..done as usize].iter().map(|&r| r as usize));
            draws.drain(..(done as usize + (rc == 1) as usize));
            if rc == 1 {                                   // the Err arm of choose_weighted (:287-290)
                let chosen = (0..n).choose(rng).unwrap() as u64;
                unsafe { ffi::spf_kmpp_push(s, chosen) };
                rows.push(chosen as usize);
            }
        }
        unsafe { ffi::spf_kmpp_free(s) };
        self.clusters = rows.into_iter().map(|r| Cluster { centroid_idx: Some(r), points: Vec::new(), depth: 0 }).collect();
    }
}
