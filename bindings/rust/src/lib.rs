//! `init_with_config` / `build` / `search` with the signatures and `Config` semantics of the reference crate
//! (src/lib.rs:118,142,183; src/core/config.rs) over libclann_b200: the whole of `ClusteredIndex::{new, build, search}`
//! (src/core/index.rs:71-91,177-289,311-439) runs behind the C ABI on the GPU. `search_batch` is the call that feeds it.
//! Source only in this repository (no Rust toolchain in the image); see INTEGRATION.md.
pub mod sys;

use std::ffi::CStr;
use thiserror::Error;

pub type Result<T> = std::result::Result<T, ClusteredIndexError>;

/// src/core/errors.rs:6-39 (the variants this path can produce).
#[derive(Debug, Error, PartialEq)]
pub enum ClusteredIndexError {
    #[error("Configuration Error: {0}")]
    ConfigError(String),
    #[error("Data Error: {0}")]
    DataError(String),
    #[error("PUFFINN Creation Error: {0}")]
    PuffinnCreationError(String),
    #[error("PUFFINN Search Error: {0}")]
    PuffinnSearchError(String),
    #[error("Index Not Found Error")]
    IndexNotFound(),
    #[error("Index Out of Bounds: {0} out of {1} length")]
    IndexOutOfBounds(usize, usize),
    #[error("Serialize Error: {0}")]
    SerializeError(String),
}

/// src/core/config.rs:18-37 without the metrics sink selector (the sqlite sink is outside the hot path).
#[derive(Debug, Clone)]
pub struct Config {
    pub num_tables: usize,
    pub num_clusters_factor: f32,
    pub k: usize,
    pub delta: f32,
    pub dataset_name: String,
}

impl Default for Config {
    fn default() -> Self {
        Self { num_tables: 10, num_clusters_factor: 1.0, k: 10, delta: 0.9, dataset_name: String::new() }
    }
}

impl Config {
    pub fn new(num_tables: usize, num_clusters_factor: f32, k: usize, delta: f32, dataset_name: &str) -> Self {
        Self { num_tables, num_clusters_factor, k, delta, dataset_name: dataset_name.to_string() }
    }
}

/// Row-major f32 rows compared by angular distance (src/metricdata/angulardata.rs); the library copies them once.
pub struct AngularData {
    pub rows: Vec<f32>,
    pub dimensions: usize,
}

impl AngularData {
    pub fn new(rows: Vec<f32>, dimensions: usize) -> Self {
        assert!(dimensions > 0 && rows.len() % dimensions == 0);
        Self { rows, dimensions }
    }
    pub fn num_points(&self) -> usize {
        self.rows.len() / self.dimensions
    }
}

pub struct ClusteredIndex {
    raw: *mut sys::clann_index,
    config: Config,
    dimensions: usize,
}

// The library serialises calls on one index internally through CUDA stream order; the handle itself is not re-entrant
// (the reference's is not either: search(&mut self)).
unsafe impl Send for ClusteredIndex {}

fn check(status: i32) -> Result<()> {
    let msg = || unsafe { CStr::from_ptr(sys::clann_last_error()) }.to_string_lossy().into_owned();
    match status {
        sys::CLANN_OK => Ok(()),
        sys::CLANN_ERR_DATA => Err(ClusteredIndexError::DataError(msg())),
        sys::CLANN_ERR_CONFIG | sys::CLANN_ERR_ARG => Err(ClusteredIndexError::ConfigError(msg())),
        sys::CLANN_ERR_CREATION | sys::CLANN_ERR_CUDA => Err(ClusteredIndexError::PuffinnCreationError(msg())),
        sys::CLANN_ERR_SEARCH => Err(ClusteredIndexError::PuffinnSearchError(msg())),
        sys::CLANN_ERR_NOT_BUILT => Err(ClusteredIndexError::IndexNotFound()),
        sys::CLANN_ERR_BOUNDS => Err(ClusteredIndexError::IndexOutOfBounds(0, 0)),
        _ => Err(ClusteredIndexError::SerializeError(msg())),
    }
}

/// src/lib.rs:118-124 -> ClusteredIndex::new (index.rs:71-91): an empty dataset is a DataError.
pub fn init_with_config(data: AngularData, config: Config) -> Result<ClusteredIndex> {
    let cfg = sys::clann_config {
        num_tables: config.num_tables as u64,
        num_clusters_factor: config.num_clusters_factor,
        k: config.k as u64,
        delta: config.delta,
    };
    let mut raw = std::ptr::null_mut();
    check(unsafe { sys::clann_init_with_config(data.rows.as_ptr(), data.num_points() as u64, data.dimensions as u32, &cfg, &mut raw) })?;
    Ok(ClusteredIndex { raw, config, dimensions: data.dimensions })
}

/// src/lib.rs:90-98: the default configuration.
pub fn init(data: AngularData) -> Result<ClusteredIndex> {
    init_with_config(data, Config::default())
}

/// src/lib.rs:142-148 -> index.rs:177-289: greedy k-center clustering + one PUFFINN-style index per cluster, on the device.
pub fn build(index: &mut ClusteredIndex) -> Result<()> {
    check(unsafe { sys::clann_build(index.raw) })
}

/// src/lib.rs:183-189 -> index.rs:311-439: up to k (distance, point id) pairs, ascending distance.
pub fn search(index: &mut ClusteredIndex, query: &[f32]) -> Result<Vec<(f32, usize)>> {
    Ok(search_batch(index, query, 1)?.pop().unwrap_or_default())
}

/// The same for `nq` queries stored row-major in `queries` — one crossing of the boundary, one GPU batch.
pub fn search_batch(index: &mut ClusteredIndex, queries: &[f32], nq: usize) -> Result<Vec<Vec<(f32, usize)>>> {
    if queries.len() != nq * index.dimensions {
        return Err(ClusteredIndexError::DataError(format!("expected {} floats, got {}", nq * index.dimensions, queries.len())));
    }
    let k = index.config.k;
    let (mut ids, mut dists, mut counts) = (vec![0u32; nq * k], vec![0f32; nq * k], vec![0u32; nq]);
    check(unsafe { sys::clann_search(index.raw, queries.as_ptr(), nq as u64, ids.as_mut_ptr(), dists.as_mut_ptr(), counts.as_mut_ptr()) })?;
    Ok((0..nq)
        .map(|q| (0..counts[q] as usize).map(|i| (dists[q * k + i], ids[q * k + i] as usize)).collect())
        .collect())
}

/// Counters the reference keeps per query (utils/metrics/mod.rs:14-20; performance.hpp:72-86), for the last batch.
pub struct QueryCounters {
    pub candidates: Vec<u64>,
    pub distance_computations: Vec<u64>,
    pub clusters_visited: Vec<u32>,
}

pub fn last_counters(index: &mut ClusteredIndex, nq: usize) -> Result<QueryCounters> {
    let mut c = QueryCounters { candidates: vec![0; nq], distance_computations: vec![0; nq], clusters_visited: vec![0; nq] };
    check(unsafe {
        sys::clann_get_counters(index.raw, nq as u64, c.candidates.as_mut_ptr(), c.distance_computations.as_mut_ptr(), c.clusters_visited.as_mut_ptr())
    })?;
    Ok(c)
}

impl Drop for ClusteredIndex {
    fn drop(&mut self) {
        unsafe { sys::clann_destroy(self.raw) }
    }
}
