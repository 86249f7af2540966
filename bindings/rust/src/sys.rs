//! Raw declarations of include/clann_b200.h (what bindgen emits with the allowlist "^clann_.*|^CPUFFINN_.*").
//! The eight CPUFFINN_* symbols keep the signatures of the reference's libpuffinn-ffi/c_binder.h:14-26, so the reference's
//! own src/puffinn_binds/puffinn_sys.rs links against this library unchanged.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_uint, c_void};

#[repr(C)]
pub struct clann_index {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct clann_config {
    pub num_tables: u64,
    pub num_clusters_factor: f32,
    pub k: u64,
    pub delta: f32,
}

pub const CLANN_OK: i32 = 0;
pub const CLANN_ERR_DATA: i32 = -1;
pub const CLANN_ERR_CONFIG: i32 = -2;
pub const CLANN_ERR_CREATION: i32 = -3;
pub const CLANN_ERR_SEARCH: i32 = -4;
pub const CLANN_ERR_NOT_BUILT: i32 = -5;
pub const CLANN_ERR_BOUNDS: i32 = -6;
pub const CLANN_ERR_SERIALIZE: i32 = -7;
pub const CLANN_ERR_CUDA: i32 = -8;
pub const CLANN_ERR_ARG: i32 = -9;

extern "C" {
    pub fn clann_init_with_config(data: *const f32, n: u64, d: u32, config: *const clann_config, out: *mut *mut clann_index) -> i32;
    pub fn clann_set_option(index: *mut clann_index, key: *const c_char, value: i64) -> i32;
    pub fn clann_build(index: *mut clann_index) -> i32;
    pub fn clann_search(index: *mut clann_index, queries: *const f32, nq: u64, ids: *mut u32, dists: *mut f32, counts: *mut u32) -> i32;
    pub fn clann_get_counters(index: *mut clann_index, nq: u64, candidates: *mut u64, distance_computations: *mut u64, clusters_visited: *mut u32) -> i32;
    pub fn clann_last_error() -> *const c_char;
    pub fn clann_destroy(index: *mut clann_index);

    #[cfg(feature = "cuda")]
    pub fn clann_search_device(index: *mut clann_index, d_queries: *const f32, nq: u64, d_ids: *mut u32, d_dists: *mut f32, d_counts: *mut u32, stream: *mut c_void) -> i32;
    #[cfg(feature = "cuda")]
    pub fn clann_search_device_async(index: *mut clann_index, d_queries: *const f32, nq: u64, d_ids: *mut u32, d_dists: *mut f32, d_counts: *mut u32) -> i32;
    #[cfg(feature = "cuda")]
    // cluster-sharded multi-GPU search (one process per GPU): NCCL transport
    pub fn clann_comm_unique_id(out: *mut u8, cap: u64) -> i32;
    pub fn clann_comm_init(index: *mut clann_index, rank: i32, world: i32, unique_id: *const u8) -> i32;
    pub fn clann_search_sharded(index: *mut clann_index, d_queries: *const f32, nq: u64, d_ids: *mut u32, d_dists: *mut f32, d_counts: *mut u32, stream: *mut c_void) -> i32;
    pub fn clann_search_sharded_submit(index: *mut clann_index, d_queries: *const f32, nq: u64, d_ids: *mut u32, d_dists: *mut f32, d_counts: *mut u32, stream: *mut c_void) -> i32;
    pub fn clann_search_sharded_flush(index: *mut clann_index, stream: *mut c_void) -> i32;
    pub fn clann_search_flush(index: *mut clann_index, stream: *mut c_void) -> i32;

    // legacy per-cluster ABI (c_binder.h:14-26)
    pub fn CPUFFINN_index_create(dataset_type: *const c_char, dataset_args: c_int) -> *mut c_void;
    pub fn CPUFFINN_index_insert_cosine(index: *mut c_void, point: *mut f32, dimension: c_int);
    pub fn CPUFFINN_index_rebuild(index: *mut c_void, num_maps: c_uint) -> u64;
    pub fn CPUFFINN_search_cosine(index: *mut c_void, query: *mut f32, k: c_uint, recall: f32, max_sim: f32, dimension: c_int) -> *mut u32;
    pub fn CPUFFINN_get_distance_computations() -> c_uint;
    pub fn CPUFFINN_clear_distance_computations();
    pub fn CPUFFINN_save_index(index: *mut c_void, file_name: *const c_char, index_id: c_int);
    pub fn CPUFFINN_load_from_file(file_name: *const c_char, dataset_name: *const c_char) -> *mut c_void;
    // puffinn::Index::search with its FilterType argument (0 Default, 1 None, 2 Simple; collection.hpp:22-34)
    pub fn clann_puffinn_search(index: *mut c_void, query: *const f32, k: u32, recall: f32, max_sim: f32, filter_type: c_int, out_ids: *mut u32, out_count: *mut u32, out_stop_depth: *mut u32) -> i32;
}
