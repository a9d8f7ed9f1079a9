//! `B200VectorStore` / `B200Backend`: the reference's own seams implemented over libcgvec_b200.so.
//!
//! * `impl VectorStore for B200VectorStore`        — codegraph-core/src/traits.rs:11-16
//! * `impl SurrealVectorBackend for B200Backend`    — codegraph-vector/src/surreal_store.rs:11-22
//!
//! Blocking FFI calls run inside `tokio::task::spawn_blocking`; errors become `CodeGraphError::Vector(msg)`
//! (codegraph-core/src/error.rs:18-19).  `search_similar(&self)` may be called concurrently (the C library pools
//! per-call scratch); `store_embeddings(&mut self)` is exclusive, exactly as the trait's receivers say.
mod ffi;

use async_trait::async_trait;
use codegraph_core::{CodeGraphError, CodeNode, NodeId, Result, VectorStore};
use codegraph_vector::SurrealVectorBackend;
use std::ffi::CStr;
use std::sync::Arc;

struct Handle(*mut ffi::cgvec_index);
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { ffi::cgvec_destroy(self.0) };
    }
}

fn last_error() -> CodeGraphError {
    let msg = unsafe { CStr::from_ptr(ffi::cgvec_last_error()) }.to_string_lossy().into_owned();
    CodeGraphError::Vector(msg)
}
fn check(rc: i32) -> Result<()> {
    if rc == ffi::CGVEC_OK { Ok(()) } else { Err(last_error()) }
}

#[derive(Clone)]
pub struct B200VectorStore {
    h: Arc<Handle>,
    dim: usize,
}

impl B200VectorStore {
    /// `device` is a CUDA ordinal; storage is f32 like `CodeNode.embedding: Option<Vec<f32>>` (node.rs:14).
    pub fn new(dimension: usize, device: i32) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::cgvec_create(dimension as u32, ffi::CGVEC_F32, &device, 1, &mut raw) })?;
        Ok(Self { h: Arc::new(Handle(raw)), dim: dimension })
    }

    /// Deployment switch: `enable_gpu` is `PerformanceConfig::enable_gpu` (config_manager.rs:362-364); `CODEGRAPH_ENABLE_GPU`
    /// overrides it and `CODEGRAPH_B200_DEVICES` ("all" | count | "0,2,5") picks the GPUs.  `Ok(None)` = switched off: keep
    /// the CPU vector store.
    pub fn from_env(dimension: usize, enable_gpu: bool) -> Result<Option<Self>> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { ffi::cgvec_create_from_env(dimension as u32, ffi::CGVEC_F32, enable_gpu as i32, &mut raw) };
        if rc == ffi::CGVEC_ERR_DISABLED { return Ok(None); }
        check(rc)?;
        Ok(Some(Self { h: Arc::new(Handle(raw)), dim: dimension }))
    }

    /// One process driving several GPUs (`cgvec_create` with n_devices > 1): rows are dealt to the devices in blocks and
    /// every search merges the per-device results over NVLink.
    pub fn with_devices(dimension: usize, devices: &[i32]) -> Result<Self> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { ffi::cgvec_create(dimension as u32, ffi::CGVEC_F32, devices.as_ptr(), devices.len() as i32, &mut raw) })?;
        Ok(Self { h: Arc::new(Handle(raw)), dim: dimension })
    }

    /// `SemanticSearch::calculate_similarity_score` for many candidates at once (search.rs:207-217, arithmetic :519-533).
    pub fn rescore(&self, query: &[f32], node_ids: &[NodeId]) -> Result<Vec<f32>> {
        let mut rows = Vec::with_capacity(node_ids.len());
        for id in node_ids {
            let mut row = 0u64;
            check(unsafe { ffi::cgvec_row_of_id(self.h.0, id.as_bytes().as_ptr(), &mut row) })?;
            rows.push(row);
        }
        let mut out = vec![0f32; rows.len()];
        check(unsafe {
            ffi::cgvec_rescore(self.h.0, query.as_ptr(), rows.as_ptr(), rows.len() as u32, ffi::CGVEC_COSINE, ffi::CGVEC_FORMULA_SEQ, out.as_mut_ptr())
        })?;
        Ok(out)
    }

    fn knn(&self, query: Vec<f32>, limit: usize) -> Result<(Vec<NodeId>, Vec<f32>)> {
        if query.is_empty() || limit == 0 {
            return Ok((Vec::new(), Vec::new())); // surreal_store.rs:62-64
        }
        if query.len() != self.dim {
            return Err(CodeGraphError::Vector(format!(
                "Query dimension {} doesn't match index dimension {}", query.len(), self.dim)));
        }
        let mut ids = vec![[0u8; 16]; limit];
        let mut scores = vec![0f32; limit];
        let mut count = 0u32;
        check(unsafe {
            ffi::cgvec_search(self.h.0, query.as_ptr(), 1, limit as u32, ffi::CGVEC_COSINE, std::ptr::null_mut(),
                              ids.as_mut_ptr(), scores.as_mut_ptr(), &mut count)
        })?;
        ids.truncate(count as usize);
        scores.truncate(count as usize);
        Ok((ids.into_iter().map(NodeId::from_bytes).collect(), scores))
    }
}

#[async_trait]
impl VectorStore for B200VectorStore {
    async fn store_embeddings(&mut self, nodes: &[CodeNode]) -> Result<()> {
        let mut ids: Vec<[u8; 16]> = Vec::new();
        let mut rows: Vec<f32> = Vec::new();
        for n in nodes {
            if let Some(e) = &n.embedding {
                if e.len() != self.dim {
                    return Err(CodeGraphError::Vector(format!(
                        "Vector dimension {} doesn't match expected {}", e.len(), self.dim)));
                }
                ids.push(*n.id.as_bytes());
                rows.extend_from_slice(e);
            }
        }
        if ids.is_empty() {
            return Ok(());
        }
        let h = self.h.clone();
        tokio::task::spawn_blocking(move || check(unsafe { ffi::cgvec_add(h.0, ids.as_ptr(), rows.as_ptr(), ids.len() as u64) }))
            .await
            .map_err(|e| CodeGraphError::Vector(e.to_string()))?
    }

    async fn search_similar(&self, query_embedding: &[f32], limit: usize) -> Result<Vec<NodeId>> {
        let this = self.clone();
        let q = query_embedding.to_vec();
        tokio::task::spawn_blocking(move || this.knn(q, limit).map(|(ids, _)| ids))
            .await
            .map_err(|e| CodeGraphError::Vector(e.to_string()))?
    }

    async fn get_embedding(&self, node_id: NodeId) -> Result<Option<Vec<f32>>> {
        let h = self.h.clone();
        let dim = self.dim;
        tokio::task::spawn_blocking(move || {
            let mut row = vec![0f32; dim];
            let rc = unsafe { ffi::cgvec_get(h.0, node_id.as_bytes().as_ptr(), row.as_mut_ptr()) };
            if rc == ffi::CGVEC_ERR_NOT_FOUND { Ok(None) } else { check(rc).map(|_| Some(row)) }
        })
        .await
        .map_err(|e| CodeGraphError::Vector(e.to_string()))?
    }
}

/// Drops in where `SurrealStorageBackend` sits (surreal_store.rs:45-53): `SurrealVectorStore::new(Arc::new(B200Backend::new(store)), ef)`.
pub struct B200Backend {
    store: Arc<tokio::sync::RwLock<B200VectorStore>>,
}

impl B200Backend {
    pub fn new(store: B200VectorStore) -> Self {
        Self { store: Arc::new(tokio::sync::RwLock::new(store)) }
    }
}

#[async_trait]
impl SurrealVectorBackend for B200Backend {
    async fn upsert_nodes(&self, nodes: &[CodeNode]) -> Result<()> {
        self.store.write().await.store_embeddings(nodes).await
    }

    /// Same shape as SurrealDbStorage::vector_search_knn (surrealdb_storage.rs:271-328): ("nodes:<uuid>", cosine
    /// distance) ascending.  `column` selects nothing (one index per dimension); `ef_search` is meaningless for an
    /// exact scan.
    async fn vector_knn(&self, _column: &str, query_embedding: Vec<f32>, limit: usize, _ef_search: usize) -> Result<Vec<(String, f32)>> {
        // The read guard travels into the blocking task and is held for the whole FFI search, so an upsert (write guard)
        // can never reallocate the device matrix under a running scan.  (The C library also serialises its write side against
        // in-flight reads with its own reader-writer lock; this keeps the shim correct on its own.)
        let store = self.store.clone().read_owned().await;
        tokio::task::spawn_blocking(move || {
            store.knn(query_embedding, limit).map(|(ids, scores)| {
                ids.into_iter().zip(scores).map(|(id, s)| (format!("nodes:{}", id), 1.0 - s)).collect()
            })
        })
        .await
        .map_err(|e| CodeGraphError::Vector(e.to_string()))?
    }

    async fn get_node_embedding(&self, node_id: NodeId) -> Result<Option<Vec<f32>>> {
        self.store.read().await.get_embedding(node_id).await
    }
}
