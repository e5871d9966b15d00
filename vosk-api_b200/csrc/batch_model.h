// batch_model.h — BatchModel: same role and public surface as the reference class
// [REF src/batch_model.h:43-66] (GetID, WaitForCompletion), rebuilt over vb::Engine.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "vb_engine.h"
#include "vb_model.h"

class BatchRecognizer;

class BatchModel {
   public:
    // The reference constructor takes no argument and reads "model/..." relative to the CWD
    // [REF src/batch_model.cc:28-37]; model_dir/options are the additive surface (include/vosk_b200.h).
    explicit BatchModel(const std::string &model_dir = "model", const std::string &options = "");
    ~BatchModel();

    uint64_t GetID(BatchRecognizer *recognizer);  // [REF src/batch_model.cc:102-104] (atomic here)
    void WaitForCompletion();                     // [REF src/batch_model.cc:118-121]

    // utterance sharding: streams are independent, so a stream lives on engine (id mod #GPUs); no collective
    vb::Engine &engine_for(uint64_t id) { return *engines_[id % engines_.size()]; }
    const vb::Model &model() const { return model_; }
    int samples_per_chunk() const { return samples_per_chunk_; }
    size_t num_engines() const { return engines_.size(); }
    vb::Engine &engine(size_t i) { return *engines_[i]; }
    std::vector<std::string> resident_results;    // texts of the last device-resident run

   private:
    vb::Model model_;
    vb::Config cfg_;
    std::vector<std::unique_ptr<vb::Engine>> engines_;  // one per GPU; streams are sharded by id (no collective)
    int samples_per_chunk_ = 0;
    std::atomic<uint64_t> last_id_{0};
};
