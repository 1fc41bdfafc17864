#include "registry.h"
#include <cstring>
#include <mutex>
#include <vector>

namespace grbda_runtime
{
    namespace
    {
        std::vector<ModelKernels *> &records()
        {
            static std::vector<ModelKernels *> r;
            return r;
        }
        std::mutex &lock()
        {
            static std::mutex m;
            return m;
        }
    } // namespace

    ModelKernels *modelRecord(uint64_t hash, const char *name, int nq, int nv, int nb, int nc)
    {
        std::lock_guard<std::mutex> g(lock());
        for (ModelKernels *k : records())
            if (k->hash == hash)
                return k;
        ModelKernels *k = new ModelKernels();
        std::memset(k, 0, sizeof(*k));
        k->hash = hash;
        k->name = name;
        k->nq = nq, k->nv = nv, k->nb = nb, k->nc = nc;
        records().push_back(k);
        return k;
    }
    const ModelKernels *findModelKernels(uint64_t hash)
    {
        std::lock_guard<std::mutex> g(lock());
        for (ModelKernels *k : records())
            if (k->hash == hash)
                return k;
        return nullptr;
    }
    int numRegisteredModels() { return (int)records().size(); }
    const ModelKernels *registeredModel(int i) { return records()[i]; }
} // namespace grbda_runtime
