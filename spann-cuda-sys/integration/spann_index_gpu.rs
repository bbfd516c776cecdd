// Goes into src/spann/spann_index.rs and src/spann/spann_builder.rs.  UNCOMPILED.
use spann_cuda_sys::{Context, Dataset, Index};

impl<const N: usize> SpannIndex<N, f32> {
    // reference spann_index.rs:56-114 — keep writing the files (unchanged readers keep working) and
    // keep the lists resident in HBM for the query path
    pub fn create_posting_lists_gpu<'a>(&mut self, ds: &Dataset<'a>, clusters: &[Cluster]) -> Result<Index<'a>, Box<dyn Error>> {
        let lists: Vec<Vec<usize>> = clusters.iter().map(|c| c.points.clone()).collect();
        let rows: Vec<usize> = clusters.iter().map(|c| c.centroid_idx.expect("centroid")).collect();
        let idx = Index::pack(ds, &lists, &rows)?;
        idx.save_dir(&self.posting_list_dir)?;           // posting_list_{id}.bin + cluster_ids.bin, bincode layout
        Ok(idx)
    }

    // reference spann_index.rs:148-197 — the single-query API forwards to the batched sibling
    pub fn find_k_nearest_neighbor_spann_gpu(&self, idx: &Index, query: &ArrayView1<f32>, k: usize) -> Option<Vec<PointData<f32>>> {
        let q = query.insert_axis(Axis(0));
        let mut res = idx.search_batch(&q.view(), k).expect("spf_search_batch");
        res.pop().flatten().map(|v| v.into_iter().map(|(point_id, vector)| PointData { point_id, vector }).collect())
    }

    pub fn find_k_nearest_neighbors_batch(&self, idx: &Index, queries: &ArrayView2<f32>, k: usize) -> Vec<Option<Vec<PointData<f32>>>> {
        idx.search_batch(queries, k).expect("spf_search_batch").into_iter()
            .map(|r| r.map(|v| v.into_iter().map(|(point_id, vector)| PointData { point_id, vector }).collect()))
            .collect()
    }
}

// reference spann_builder.rs:66-75 — the dense centroid matrix comes from the centroids.bin sidecar
// written at build time (u64 rows, u64 cols, f32 data); output.kdtree stays for unmodified readers
pub fn load_gpu<'a>(ctx: &'a Context, output_path: &str) -> Result<Index<'a>, Box<dyn Error>> {
    let raw = std::fs::read(format!("{output_path}/centroids.bin"))?;
    let rows = u64::from_le_bytes(raw[0..8].try_into()?) as usize;
    let cols = u64::from_le_bytes(raw[8..16].try_into()?) as usize;
    let data: Vec<f32> = raw[16..].chunks_exact(4).map(|b| f32::from_le_bytes(b.try_into().unwrap())).collect();
    let cen = ndarray::Array2::from_shape_vec((rows, cols), data)?;
    Ok(Index::load_dir(ctx, output_path, &cen.view())?)
}
