//! `extern "C"` view of include/neurokmer.h.  NOT compiled in this repository's image (no rustc).
//! Each line names the reference item it replaces (paths relative to the reference crate).
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct NkCounter {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct NkConfig {
    pub k: u32,
    pub refractory: u32,
    pub pool_size: u64,
    pub steps: u64,
    pub spike_cost: f64,
    pub threshold: f32,
    pub leak: f32,
    pub use_canonical: i32,
    pub device: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct NkTopEntry {
    pub idx: u64,
    pub spikes: u64,
    pub uniques: u32,
    pub _pad: u32,
}
pub const NK_UNIQUES_NOT_COMPUTED: u32 = 0xFFFF_FFFF;

extern "C" {
    pub fn nk_config_default(cfg: *mut NkConfig) -> c_int;
    pub fn nk_create(cfg: *const NkConfig, out: *mut *mut NkCounter) -> c_int; // SpikingKmerCounter::new   src/spiking_hash.rs:40-77
    // ONE input sharded over several GPUs of this process (the reference's rayon fold/reduce over cores, src/spiking_hash.rs:94-154)
    pub fn nk_device_count(n: *mut i32) -> c_int;
    pub fn nk_create_multi(cfg: *const NkConfig, devices: *const i32, n_devices: i32, out: *mut *mut NkCounter) -> c_int;
    pub fn nk_destroy(h: *mut NkCounter) -> c_int; // Drop
    pub fn nk_reset(h: *mut NkCounter) -> c_int;
    pub fn nk_last_error() -> *const c_char;
    pub fn nk_set_steps(h: *mut NkCounter, steps: u64) -> c_int; // set_steps :688-690
    pub fn nk_get_steps(h: *const NkCounter, steps: *mut u64) -> c_int; // get_steps :693-695
    pub fn nk_process_batch(h: *mut NkCounter, bases: *const u8, offsets: *const u64, nseq: u64) -> c_int; // process_parallel :84-201
    pub fn nk_process_sequence(h: *mut NkCounter, seq: *const u8, len: u64) -> c_int; // process_sequence :203-273
    pub fn nk_process_file(h: *mut NkCounter, path: *const c_char, streaming: c_int) -> c_int; // process_file_streaming :277-486
    pub fn nk_stream_begin(h: *mut NkCounter) -> c_int;
    pub fn nk_stream_push(h: *mut NkCounter, bases: *const u8, offsets: *const u64, nseq: u64) -> c_int;
    pub fn nk_stream_end(h: *mut NkCounter) -> c_int;
    pub fn nk_simulate(h: *mut NkCounter) -> c_int; // simulate_spikes_auto :697-714
    pub fn nk_top_n(h: *mut NkCounter, top_n: u64, out: *mut NkTopEntry, n_out: *mut u64) -> c_int; // top_abundant_neurons :661-673
    pub fn nk_total_spikes(h: *const NkCounter, out: *mut u64) -> c_int; // energy.total_spikes()  src/models.rs:166-168
    pub fn nk_energy_used(h: *const NkCounter, out: *mut f64) -> c_int; // energy_used :684-686
    pub fn nk_enable_exact_counts(h: *mut NkCounter, on: c_int) -> c_int; // counts / kmer_per_neuron :26-27
    pub fn nk_get_count(h: *mut NkCounter, kmer: u64, count: *mut u32, found: *mut i32) -> c_int; // get_count :675-678
    pub fn nk_set_file_uniques(h: *mut NkCounter, top_n: u64) -> c_int; // `uniques` of the printed rows, src/main.rs:54-61
    pub fn nk_uniques_begin(h: *mut NkCounter, top_n: u64) -> c_int;
    pub fn nk_uniques_push(h: *mut NkCounter, bases: *const u8, offsets: *const u64, nseq: u64) -> c_int;
    pub fn nk_uniques_end(h: *mut NkCounter) -> c_int;
    pub fn nk_pack_kmer(kmer: *const u8, len: u64) -> u64; // utils::pack_kmer  src/utils.rs:26-39
    // pre-packed input (2 bits per base + `other` bits)
    pub fn nk_packed_code_words(nbases: u64) -> u64;
    pub fn nk_packed_other_words(nbases: u64) -> u64;
    pub fn nk_pack_bases(bases: *const u8, nbases: u64, codes: *mut u32, other: *mut u32, threads: c_int, n_other: *mut u64) -> c_int;
    pub fn nk_process_batch_packed(h: *mut NkCounter, codes: *const u32, other: *const u32, offsets: *const u64, nseq: u64) -> c_int;
    pub fn nk_stream_push_packed(h: *mut NkCounter, codes: *const u32, other: *const u32, offsets: *const u64, nseq: u64) -> c_int;
    // multi-GPU (one process per GPU)
    pub fn nk_stream_accumulated(h: *mut NkCounter, dev_currents: *mut *mut c_void) -> c_int;
    pub fn nk_stream_finish(h: *mut NkCounter) -> c_int;
    pub fn nk_dist_export(h: *mut NkCounter, ipc_handle_64_bytes: *mut c_void, raw_acc: *mut *mut c_void) -> c_int;
    pub fn nk_dist_setup(h: *mut NkCounter, rank: c_int, world: c_int, ipc_handles: *const c_void, raw_ptrs: *const *mut c_void) -> c_int;
    pub fn nk_dist_run(h: *mut NkCounter) -> c_int;
    // pinned host memory (batches in it are read in place by the count kernel)
    pub fn nk_host_alloc(ptr: *mut *mut c_void, nbytes: u64) -> c_int;
    pub fn nk_host_free(ptr: *mut c_void) -> c_int;
}
