//! spann-cuda-sys — `extern "C"` declarations (module `ffi`, generated from
//! `include/spfresh_b200.h` by `tools/gen_rust_ffi.py`) and thin RAII wrappers.
//!
//! UNCOMPILED: the authoring image has no Rust toolchain.  The same ABI is exercised by the ctypes
//! binding (`spfresh_b200/_capi.py`) and by the C++ host layer (`host/spfresh.hpp`), both of which
//! run in the test suite; `tests/test_abi.py` checks that `ffi.rs` declares exactly the exported
//! symbols.
//!
//! Conventions: every call returns 0 or a negative `SPF_E_*`; the message is
//! thread-local (`spf_last_error`).  `usize` <-> `u64`, cluster slots are `u32`.  Host pointers are
//! only read during the call.  A context may be shared between threads (calls are serialised).
pub mod ffi;

use std::ffi::{CStr, CString};
use std::fmt;
use std::ptr;

use ndarray::ArrayView2;

#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl fmt::Display for Error {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "libspfresh_b200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}
pub type Result<T> = std::result::Result<T, Error>;

fn check(rc: i32) -> Result<i32> {
    if rc < 0 {
        let message = unsafe { CStr::from_ptr(ffi::spf_last_error()) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message })
    } else {
        Ok(rc)
    }
}

/// Metric ids of `Config::to_clustering_params` (src/spann/config.rs:92-100).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum MetricKind {
    Euclidean = 0,
    Manhattan = 1,
    Chebyshev = 2,
}

/// One B200 + its stream.  `Send + Sync`: the library serialises calls on a context.
pub struct Context(*mut ffi::spf_ctx);
unsafe impl Send for Context {}
unsafe impl Sync for Context {}
impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::spf_ctx_create(device, &mut h) })?;
        Ok(Context(h))
    }
    pub fn raw(&self) -> *mut ffi::spf_ctx {
        self.0
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::spf_ctx_destroy(self.0) }
    }
}

/// The borrowed `ArrayView2<f32>` of `HierarchicalClustering` / `SpannIndexBuilder`
/// (hierarchical.rs:45, spann_builder.rs:10,20) copied to HBM.
pub struct Dataset<'c> {
    h: *mut ffi::spf_dataset,
    pub n: usize,
    pub d: usize,
    _ctx: &'c Context,
}
impl<'c> Dataset<'c> {
    pub fn upload(ctx: &'c Context, data: &ArrayView2<f32>) -> Result<Self> {
        assert!(data.strides()[1] == 1, "row-major rows expected (the reference calls as_slice() on rows)");
        let (n, d) = data.dim();
        let mut h = ptr::null_mut();
        check(unsafe {
            ffi::spf_dataset_upload(ctx.raw(), data.as_ptr(), n as u64, d as u32, data.strides()[0] as u64, &mut h)
        })?;
        Ok(Dataset { h, n, d, _ctx: ctx })
    }
    pub fn raw(&self) -> *mut ffi::spf_dataset {
        self.h
    }

    /// `assign_points_to_clusters` (hierarchical.rs:295-364): per cluster the member rows, in input order.
    pub fn assign(&self, metric: MetricKind, point_indices: Option<&[usize]>, centroid_rows: &[usize],
                  boundary: f32) -> Result<Assignment> {
        let pidx: Option<Vec<u64>> = point_indices.map(|p| p.iter().map(|&x| x as u64).collect());
        let crow: Vec<u64> = centroid_rows.iter().map(|&x| x as u64).collect();
        let (pp, m) = match &pidx {
            Some(v) => (v.as_ptr(), v.len() as u64),
            None => (ptr::null(), self.n as u64),
        };
        let mut r = ptr::null_mut();
        check(unsafe {
            ffi::spf_assign(self.h, metric as i32, pp, m, crow.as_ptr(), crow.len() as u32, boundary, 0, &mut r)
        })?;
        Ok(Assignment { h: r })
    }

    /// `update_centroids` (hierarchical.rs:138-181) on the assignment just computed.
    pub fn update_medoids(&self, metric: MetricKind, a: &Assignment, old_rows: &[usize]) -> Result<Vec<usize>> {
        let old: Vec<u64> = old_rows.iter().map(|&x| x as u64).collect();
        let mut new = vec![0u64; old.len()];
        check(unsafe {
            ffi::spf_update_medoids_from(self.h, metric as i32, a.h, old.as_ptr(), new.as_mut_ptr(), ptr::null_mut())
        })?;
        Ok(new.into_iter().map(|x| x as usize).collect())
    }

    /// The fold of `create_subclusters` (hierarchical.rs:112-126).
    pub fn farthest(&self, metric: MetricKind, c1: usize, members: &[usize]) -> Result<usize> {
        let mem: Vec<u64> = members.iter().map(|&x| x as u64).collect();
        let mut out = 0u64;
        check(unsafe { ffi::spf_farthest(self.h, metric as i32, c1 as u64, mem.as_ptr(), mem.len() as u64, &mut out) })?;
        Ok(out as usize)
    }
}
impl Drop for Dataset<'_> {
    fn drop(&mut self) {
        unsafe { ffi::spf_dataset_free(self.h) }
    }
}

/// Result of one assign call, resident on the device until fetched.
pub struct Assignment {
    h: *mut ffi::spf_assign_result,
}
impl Assignment {
    pub fn raw(&self) -> *const ffi::spf_assign_result {
        self.h
    }
    /// The `Vec<Vec<usize>>` the reference returns (hierarchical.rs:353-363).
    pub fn cluster_lists(&self) -> Result<Vec<Vec<usize>>> {
        let k = unsafe { ffi::spf_assign_clusters(self.h) } as usize;
        let total = unsafe { ffi::spf_assign_total(self.h) } as usize;
        let mut offsets = vec![0u64; k + 1];
        let mut members = vec![0u64; total];
        check(unsafe {
            ffi::spf_assign_fetch(self.h, ptr::null_mut(), ptr::null_mut(), offsets.as_mut_ptr(), members.as_mut_ptr())
        })?;
        Ok((0..k)
            .map(|c| members[offsets[c] as usize..offsets[c + 1] as usize].iter().map(|&x| x as usize).collect())
            .collect())
    }
}
impl Drop for Assignment {
    fn drop(&mut self) {
        unsafe { ffi::spf_assign_free(self.h) }
    }
}

/// Posting lists + centroids resident in HBM (stands in for `FileBasedPostingListStore` + the kd-tree).
pub struct Index<'c> {
    h: *mut ffi::spf_index,
    pub d: usize,
    _ctx: &'c Context,
}
impl<'c> Index<'c> {
    /// `create_posting_lists` + `build_kdtree` (spann_index.rs:56-114).
    pub fn pack(ds: &Dataset<'c>, clusters: &[Vec<usize>], centroid_rows: &[usize]) -> Result<Self> {
        let mut offsets = vec![0u64; clusters.len() + 1];
        for (i, c) in clusters.iter().enumerate() {
            offsets[i + 1] = offsets[i] + c.len() as u64;
        }
        let members: Vec<u64> = clusters.iter().flatten().map(|&x| x as u64).collect();
        let crow: Vec<u64> = centroid_rows.iter().map(|&x| x as u64).collect();
        let mut h = ptr::null_mut();
        let n = clusters.len() as u32;
        check(unsafe { ffi::spf_index_pack(ds.raw(), offsets.as_ptr(), members.as_ptr(), crow.as_ptr(), n, 0, n, &mut h) })?;
        Ok(Index { h, d: ds.d, _ctx: ds._ctx })
    }
    /// `SpannIndexBuilder::load` (spann_builder.rs:66-75): the reference's files + the dense centroid matrix.
    pub fn load_dir(ctx: &'c Context, dir: &str, centroids: &ArrayView2<f32>) -> Result<Self> {
        let (nlists, d) = centroids.dim();
        let cdir = CString::new(dir).expect("path contains NUL");
        let cen = centroids.as_standard_layout();
        let mut h = ptr::null_mut();
        check(unsafe { ffi::spf_index_load_dir(ctx.raw(), cdir.as_ptr(), cen.as_ptr(), nlists as u32, d as u32, &mut h) })?;
        Ok(Index { h, d, _ctx: ctx })
    }
    pub fn save_dir(&self, dir: &str) -> Result<()> {
        let cdir = CString::new(dir).expect("path contains NUL");
        check(unsafe { ffi::spf_index_save_dir(self.h, cdir.as_ptr()) }).map(|_| ())
    }
    /// Batched `find_k_nearest_neighbor_spann` (spann_index.rs:148-197): per query `None` or the
    /// (point_id, vector) list the reference returns.
    pub fn search_batch(&self, queries: &ArrayView2<f32>, k: usize) -> Result<Vec<Option<Vec<(usize, Vec<f32>)>>>> {
        let (nq, d) = queries.dim();
        assert_eq!(d, self.d, "Query length mismatch");
        let q = queries.as_standard_layout();
        let mut ids = vec![0u64; nq * k];
        let mut dists = vec![0f32; nq * k];
        let mut counts = vec![0u32; nq];
        let mut vectors = vec![0f32; nq * k * d];
        check(unsafe {
            ffi::spf_search_batch(self.h, q.as_ptr(), nq as u64, k as u32, 0, 1.2, ids.as_mut_ptr(), dists.as_mut_ptr(),
                                  counts.as_mut_ptr(), vectors.as_mut_ptr(), ptr::null_mut())
        })?;
        Ok((0..nq)
            .map(|i| {
                let n = counts[i] as usize;
                if n == 0 {
                    None                                  // spann_index.rs:183-186
                } else {
                    Some((0..n).map(|j| (ids[i * k + j] as usize, vectors[(i * k + j) * d..(i * k + j + 1) * d].to_vec())).collect())
                }
            })
            .collect())
    }
}
impl Drop for Index<'_> {
    fn drop(&mut self) {
        unsafe { ffi::spf_index_free(self.h) }
    }
}
