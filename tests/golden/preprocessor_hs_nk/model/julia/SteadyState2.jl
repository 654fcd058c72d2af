function steady_state!(ys_::Vector{<: Real}, exo_::Vector{<: Real}, params::Vector{<: Real})
@inbounds begin
    ys_[1]=0;
    ys_[3]=0;
    ys_[2]=0;
    ys_[5]=0;
    ys_[4]=0;
    ys_[6]=params[9];
    ys_[7]=params[8];
    ys_[8]=params[10] + params[8] + 4*params[9];
end
end
