//! Drop-in replacement of the reference's src/spiking_hash.rs: the same public type and methods, every call
//! forwarded to libneurokmer.so.  NOT compiled in this repository's image (no rustc).
use crate::ffi;
use crate::NeuroResult;

pub struct SpikingKmerCounter {
    h: *mut ffi::NkCounter,
    pub k: usize,
    pub use_canonical: bool,
    pool_size: usize,
}
unsafe impl Send for SpikingKmerCounter {} // one thread at a time per handle (&mut self), like today

fn check(rc: i32) -> NeuroResult<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(ffi::nk_last_error()) }.to_string_lossy().into_owned();
    Err(format!("libneurokmer error {rc}: {msg}").into())
}

impl SpikingKmerCounter {
    pub fn new(k: usize, threshold: f32, leak: f32, refractory: u32, spike_cost: f64, pool_size: usize,
               use_canonical: bool) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::NkConfig>() };
        unsafe { ffi::nk_config_default(&mut cfg) };
        cfg.k = k as u32;
        cfg.threshold = threshold;
        cfg.leak = leak;
        cfg.refractory = refractory;
        cfg.spike_cost = spike_cost;
        cfg.pool_size = pool_size as u64;
        cfg.use_canonical = use_canonical as i32;
        let mut h = std::ptr::null_mut();
        // the reference's `new` is infallible; a missing B200 is a hard error here (there is no CPU fallback)
        check(unsafe { ffi::nk_create(&cfg, &mut h) }).expect("nk_create");
        Self { h, k, use_canonical, pool_size }
    }

    /// The same counter with every input sharded over `gpus` GPUs of this process (0 = all of them): batches are
    /// cut by window start with a k-1 overlap, the per-neuron sums meet over NVLink peer memory.  Same results.
    pub fn new_multi(k: usize, threshold: f32, leak: f32, refractory: u32, spike_cost: f64, pool_size: usize,
                     use_canonical: bool, gpus: usize) -> Self {
        let mut cfg = unsafe { std::mem::zeroed::<ffi::NkConfig>() };
        unsafe { ffi::nk_config_default(&mut cfg) };
        cfg.k = k as u32;
        cfg.threshold = threshold;
        cfg.leak = leak;
        cfg.refractory = refractory;
        cfg.spike_cost = spike_cost;
        cfg.pool_size = pool_size as u64;
        cfg.use_canonical = use_canonical as i32;
        let mut n = gpus as i32;
        if n == 0 {
            unsafe { ffi::nk_device_count(&mut n) };
        }
        let mut h = std::ptr::null_mut();
        check(unsafe { ffi::nk_create_multi(&cfg, std::ptr::null(), n.max(1), &mut h) }).expect("nk_create_multi");
        Self { h, k, use_canonical, pool_size }
    }

    /// `&[Vec<u8>]` is flattened into one byte array + offsets (the ABI's batch form); a single sequence is
    /// handed over where it lies (no copy).  Pageable memory is staged by the library's pool of host threads.
    pub fn process_parallel(&mut self, seqs: &[Vec<u8>]) {
        if seqs.len() == 1 {
            let offsets = [0u64, seqs[0].len() as u64];
            check(unsafe { ffi::nk_process_batch(self.h, seqs[0].as_ptr(), offsets.as_ptr(), 1) }).expect("nk_process_batch");
            return;
        }
        let mut offsets = Vec::with_capacity(seqs.len() + 1);
        let mut bases = Vec::with_capacity(seqs.iter().map(Vec::len).sum());
        offsets.push(0u64);
        for s in seqs {
            bases.extend_from_slice(s);
            offsets.push(bases.len() as u64);
        }
        check(unsafe { ffi::nk_process_batch(self.h, bases.as_ptr(), offsets.as_ptr(), seqs.len() as u64) })
            .expect("nk_process_batch");
    }

    /// Pre-packed batch: `codes` / `other` as nk_pack_bases lays them out (or the caller's own packer).
    pub fn process_parallel_packed(&mut self, codes: &[u32], other: Option<&[u32]>, offsets: &[u64]) {
        let o = other.map_or(std::ptr::null(), |x| x.as_ptr());
        check(unsafe { ffi::nk_process_batch_packed(self.h, codes.as_ptr(), o, offsets.as_ptr(), (offsets.len() - 1) as u64) })
            .expect("nk_process_batch_packed");
    }

    pub fn process_sequence(&mut self, seq: &[u8]) {
        check(unsafe { ffi::nk_process_sequence(self.h, seq.as_ptr(), seq.len() as u64) }).expect("nk_process_sequence");
    }

    pub fn process_file_streaming(&mut self, path: &str) -> NeuroResult<()> {
        let c = std::ffi::CString::new(path)?;
        check(unsafe { ffi::nk_process_file(self.h, c.as_ptr(), 1) }) // NK_ERR_IO <=> the reference's Err
    }

    /// main.rs:44-45: stream_sequences().collect() + process_parallel
    pub fn process_file_in_memory(&mut self, path: &str) -> NeuroResult<()> {
        let c = std::ffi::CString::new(path)?;
        check(unsafe { ffi::nk_process_file(self.h, c.as_ptr(), 0) })
    }

    /// Ask nk_process_file to fill the `uniques` column of the top `n` rows (second read of the file).
    pub fn want_uniques(&mut self, n: usize) {
        unsafe { ffi::nk_set_file_uniques(self.h, n as u64) };
    }

    pub fn top_abundant_neurons(&self, top_n: usize) -> Vec<(usize, u64, u32)> {
        let n = top_n.min(self.pool_size);
        let mut out = vec![ffi::NkTopEntry::default(); n];
        let mut got = 0u64;
        check(unsafe { ffi::nk_top_n(self.h, n as u64, out.as_mut_ptr(), &mut got) }).expect("nk_top_n");
        out.truncate(got as usize);
        // `uniques` is NK_UNIQUES_NOT_COMPUTED unless the exact table or the uniques pass ran
        out.into_iter().map(|e| (e.idx as usize, e.spikes, e.uniques)).collect()
    }

    pub fn get_count(&self, kmer: u64) -> Option<u32> {
        let (mut c, mut f) = (0u32, 0i32);
        let rc = unsafe { ffi::nk_get_count(self.h, kmer, &mut c, &mut f) };
        if rc == 0 && f != 0 { Some(c) } else { None }
    }

    pub fn energy_used(&self) -> f64 {
        let mut v = 0.0;
        unsafe { ffi::nk_energy_used(self.h, &mut v) };
        v
    }
    pub fn total_spikes(&self) -> u64 {
        let mut v = 0;
        unsafe { ffi::nk_total_spikes(self.h, &mut v) };
        v
    }
    pub fn set_steps(&mut self, steps: usize) {
        unsafe { ffi::nk_set_steps(self.h, steps as u64) };
    }
    pub fn get_steps(&self) -> usize {
        let mut v = 0;
        unsafe { ffi::nk_get_steps(self.h, &mut v) };
        v as usize
    }
    pub fn simulate_spikes_auto(&mut self) {
        check(unsafe { ffi::nk_simulate(self.h) }).expect("nk_simulate");
    }
}

impl Drop for SpikingKmerCounter {
    fn drop(&mut self) {
        unsafe { ffi::nk_destroy(self.h) };
    }
}
