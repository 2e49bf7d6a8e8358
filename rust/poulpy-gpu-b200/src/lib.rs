//! `poulpy-gpu-b200`: B200 (sm_100a) backend of `poulpy-hal` over the C ABI of `libpoulpy_b200.so`.
//!
//! This file is the maintainer-side half of the drop-in boundary described in INTEGRATION.md.  It could not be compiled in the image this
//! repository was built in (no `cargo` / `rustc`), so it is a specification: every `ffi::pgb_*` symbol it uses exists in the library
//! (`tests/test_abi.py`), `ffi.rs` is generated from the header, and the Python binding `poulpy_b200/hal.py` exercises the identical call
//! sequences in the parity tests.  Trait and method names are those of the reference (poulpy-hal/src/oep/hal_impl.rs:25-755,
//! poulpy-core/src/oep/core_impl.rs:34-130).
#![feature(trait_alias)]
pub mod ffi;

use std::ptr::NonNull;

use poulpy_hal::layouts::{Backend, Module};

/// NTT120 flavour: ScalarPrep = 4 x u32 canonical residues (16 B), ScalarBig = i128.
pub struct B200Ntt120;
/// FFT64 flavour: ScalarPrep = f64, ScalarBig = i64.
pub struct B200Fft64;

/// CUDA managed memory from `pgb_alloc_bytes`: host-dereferenceable (`DataRef: AsRef<[u8]>`, poulpy-hal/src/layouts/mod.rs:56) and
/// device-accessible; freed with `pgb_free`.
pub struct ManagedBuf {
    ptr: NonNull<u8>,
    len: usize,
}
impl ManagedBuf {
    fn new(len: usize) -> Self {
        let p = unsafe { ffi::pgb_alloc_bytes(len) } as *mut u8;
        ManagedBuf { ptr: NonNull::new(p).expect("pgb_alloc_bytes failed"), len }
    }
}
impl AsRef<[u8]> for ManagedBuf {
    fn as_ref(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr.as_ptr(), self.len) }
    }
}
impl AsMut<[u8]> for ManagedBuf {
    fn as_mut(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr.as_ptr(), self.len) }
    }
}
impl Drop for ManagedBuf {
    fn drop(&mut self) {
        unsafe { ffi::pgb_free(self.ptr.as_ptr() as *mut _) }
    }
}
unsafe impl Send for ManagedBuf {}
unsafe impl Sync for ManagedBuf {}

/// 16-byte ScalarPrep of the NTT120 flavour (opaque to every caller: only `*_prepare` / `dft_apply` write it).
#[repr(C)]
#[derive(Copy, Clone, Default, Debug, PartialEq, bytemuck::Zeroable, bytemuck::Pod)]
pub struct Res4x32(pub [u32; 4]);

impl Backend for B200Ntt120 {
    type ScalarPrep = Res4x32;
    type ScalarBig = i128;
    type OwnedBuf = ManagedBuf;
    type Handle = ffi::PgbModule;
    fn alloc_bytes(len: usize) -> ManagedBuf {
        ManagedBuf::new(len)
    }
    fn from_bytes(bytes: Vec<u8>) -> ManagedBuf {
        let mut b = ManagedBuf::new(bytes.len());
        b.as_mut().copy_from_slice(&bytes);
        b
    }
    unsafe fn destroy(handle: NonNull<Self::Handle>) {
        unsafe { ffi::pgb_module_destroy(handle.as_ptr()) }
    }
}
impl Backend for B200Fft64 {
    type ScalarPrep = f64;
    type ScalarBig = i64;
    type OwnedBuf = ManagedBuf;
    type Handle = ffi::PgbModule;
    fn alloc_bytes(len: usize) -> ManagedBuf {
        ManagedBuf::new(len)
    }
    fn from_bytes(bytes: Vec<u8>) -> ManagedBuf {
        let mut b = ManagedBuf::new(bytes.len());
        b.as_mut().copy_from_slice(&bytes);
        b
    }
    unsafe fn destroy(handle: NonNull<Self::Handle>) {
        unsafe { ffi::pgb_module_destroy(handle.as_ptr()) }
    }
}

/// `HalImpl::new` (hal_impl.rs:320) for either flavour.
pub fn module_new<B: Backend<Handle = ffi::PgbModule>>(n: u64, flavour: i32, device: i32) -> Module<B> {
    let mut h: *mut ffi::PgbModule = std::ptr::null_mut();
    ffi::check(unsafe { ffi::pgb_module_new(n, flavour, device, &mut h) });
    unsafe { Module::from_nonnull(NonNull::new(h).unwrap(), n) }
}

/// View of a `VecZnx` / `VecZnxDft` / `VecZnxBig` for the C ABI (the five `#[repr(C)]` fields of the reference's layouts).
pub fn view(data: *const u8, n: usize, cols: usize, size: usize, max_size: usize) -> ffi::PgbVecZnx {
    ffi::PgbVecZnx { data: data as *mut _, n: n as u64, cols: cols as u64, size: size as u64, max_size: max_size as u64 }
}

// The `unsafe impl HalImpl<B200Ntt120> for B200Ntt120` block is one forward per method, e.g.
//
//     fn vec_znx_dft_apply<R, A>(module: &Module<Self>, step: usize, offset: usize, res: &mut R, res_col: usize, a: &A, a_col: usize)
//     where R: VecZnxDftToMut<Self>, A: VecZnxToRef {                                                  // hal_impl.rs:529
//         let (mut r, a) = (res.to_mut(), a.to_ref());
//         let (mut rv, av) = (view(r.data.as_ptr(), r.n(), r.cols(), r.size(), r.max_size()), view(a.data.as_ptr(), a.n(), a.cols(), a.size(), a.max_size()));
//         ffi::check(unsafe { ffi::pgb_vec_znx_dft_apply(module.ptr() as *mut _, step as u64, offset as u64, &mut rv, res_col as u64, &av, a_col as u64) });
//     }
//
// and the `CoreImpl` overrides (core_impl.rs:36-52, :114-130) forward `glwe_keyswitch` / `glwe_external_product` to
// `ffi::pgb_glwe_keyswitch_batched` / `ffi::pgb_glwe_external_product_batched` with `PgbBatch { count: 1, .. }` followed by
// `ffi::pgb_module_sync`; INTEGRATION.md sections 3-4 list every method and the entry it lands on.  A `GGLWEPrepared` / `GGSWPrepared`
// that lives as long as its Rust value is immutable: its constructor calls `ffi::pgb_gadget_key_pin`, its `Drop` `ffi::pgb_gadget_key_unpin`.
