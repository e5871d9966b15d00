// batch_recognizer.h — BatchRecognizer: same public surface as the reference class
// [REF src/batch_recognizer.h:28-53], rebuilt over vb::Engine streams.
#pragma once
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <vector>

#include "batch_model.h"
#include "vb_result.h"

class BatchRecognizer {
   public:
    BatchRecognizer(BatchModel *model, float sample_frequency);
    ~BatchRecognizer();

    void AcceptWaveform(const char *data, int len);  // [REF src/batch_recognizer.cc:115-181]
    int GetNumPendingChunks();                       // [REF :199-202]
    const char *FrontResult();                       // [REF :183-189]
    void Pop();                                      // [REF :191-197]
    void FinishStream();                             // [REF :37-41]
    void SetNLSML(bool nlsml);                       // [REF :109-112]

    const char *PartialResult();                     // additive: best path so far as {"partial" : "..."} [REF src/recognizer.cc:795-802]
    int PartialFrames();                             // additive: decoder frames (30 ms) the partial covers
    void EnableCapture();                            // additive: test taps (include/vosk_b200.h)
    vb::Capture *capture() { return stream_->capture.get(); }

   private:
    // results are produced on the engine worker thread and consumed on the caller's thread: unlike the
    // reference's unguarded std::queue [REF src/batch_recognizer.h:50] this one is locked, and it is
    // shared so a result arriving after the recognizer was freed has somewhere safe to land.
    struct Sink {
        std::mutex mu;
        std::queue<std::string> results;
        std::map<int, std::string> early;  // segments finished out of order by the lattice pool wait here for their turn
        int next_seq = 0;
        bool nlsml = false;
    };
    BatchModel *model_;
    uint64_t id_;
    float sample_frequency_;
    std::shared_ptr<Sink> sink_;
    std::shared_ptr<vb::Stream> stream_;
    vb::LinearResampler resampler_;
    std::vector<int16_t> buffer_;
    // device-side resampling: the calls whose 16 kHz output has not been handed over yet (raw input-rate samples)
    struct Call {
        std::vector<int16_t> in;
        long long n_out = 0, taken = 0;
    };
    std::deque<Call> calls_;
    long long avail_ = 0;  // 16 kHz samples the queued calls still hold
    bool device_resample_ = false;
    void push_resampled_chunk(int n_out, bool last);
    std::string front_;  // keeps the string returned by FrontResult alive until Pop
    std::string partial_;
    bool finished_ = false;
};
